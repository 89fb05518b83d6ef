"""CPU bit-plane oracle for the QGTC hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the arithmetic of the reference kernels in
``/root/reference/QGTC_module`` (kernel.h / QGTC_device.cu).  It is the checker
used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
leg.  Nothing under ``qgtc_ppopp22_b200/`` may import it: the product path is the
CUDA library and must fail loudly when that library is missing.

Parity status: the reference ships no asserting tests and no golden vectors
(SURVEY.md section 4), so this oracle is pinned two ways:
  * hand-derived known answers in tests/test_oracle.py, and
  * outputs of the UNMODIFIED reference kernels, rebuilt for sm_100a by
    ``oracle/build_ref.py`` and executed on a B200 by
    ``tests/golden/make_golden_from_reference.py``; the resulting vectors are
    committed under ``tests/golden/`` and replayed on CPU by tests/test_golden.py.
The vectors committed under tests/golden/ (9 cases, generated on an NVIDIA B200, see
tests/golden/report_b200.json) are reproduced bit for bit by this file and by oracle.c, so parity
is pinned by the reference's own outputs on its well-defined domain.  (The UNMODIFIED build of the
reference mis-maps tiles on a B200 -- it reads %warpid -- and disagrees with itself at real sizes;
the vectors come from the logical-warp-index build, see oracle/ref_logical_warpid.h.)

All packed tensors are handled as flat ``uint32`` word arrays ("raw memory"),
because the reference's drivers chain ops with mismatched layouts and parity is
defined per op on raw memory.  Reads outside an array
return 0 and writes outside are dropped ("zero-extension", SURVEY.md 8c rule 10):
that is this repo's documented rule for the reference's out-of-bounds corners.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# padding macros: utility.h:33-45
# --------------------------------------------------------------------------------------


def S8(x: int) -> int:
    return (x + 7) >> 3


def S128(x: int) -> int:
    return (x + 127) >> 7


def P8(x: int) -> int:
    return S8(x) << 3


def P128(x: int) -> int:
    return S128(x) << 7


LAYOUT_ROW = 0     # val2bit(col_major=False)            kernel.h:204-242
LAYOUT_COL128 = 1  # val2bit(col_major=True, output_layer=False)  kernel.h:75-106
LAYOUT_COL8 = 2    # val2bit(col_major=True, output_layer=True)   QGTC_device.cu:83


def packed_shape(layout: int, bits: int, H: int, W: int) -> tuple[int, int]:
    """Reported tensor shape of a packed tensor (QGTC_device.cu:83,97,115)."""
    if layout == LAYOUT_ROW:
        return (bits * P8(H), S128(W) * 4)
    if layout == LAYOUT_COL128:
        return (bits * S128(H) * 4, P128(W))
    if layout == LAYOUT_COL8:
        return (bits * S128(H) * 4, P8(W))
    raise ValueError(layout)


def packed_words(layout: int, bits: int, H: int, W: int) -> int:
    a, b = packed_shape(layout, bits, H, W)
    return a * b


# --------------------------------------------------------------------------------------
# quantize: kernel.h:39-44 (clip) and :49-71 (Quantize_val)
# --------------------------------------------------------------------------------------


def quantize_val(x: np.ndarray, nbits: int) -> np.ndarray:
    """q = __float2int_rn(clip(x, 0, 2^nbits)); clip: x<0 -> 1, x>2^b -> 2^b-1 (strict)."""
    x = np.asarray(x, dtype=np.float32)
    ub = np.float32(1 << nbits)
    y = np.where(x < 0, np.float32(1.0), np.where(x > ub, ub - np.float32(1.0), x)).astype(np.float32)
    nan = np.isnan(y)
    q = np.rint(np.where(nan, np.float32(0), y)).astype(np.int64)   # rint == round-half-even; NaN -> 0
    return q.astype(np.int32)


# --------------------------------------------------------------------------------------
# bit helpers (MSB-first words: element e of a 32-group sits in bit 31-(e%32); kernel.h:98,234)
# --------------------------------------------------------------------------------------


def _pack_msb_first(bits01: np.ndarray) -> np.ndarray:
    """[..., 32*n] array of 0/1 -> [..., n] uint32, element e -> bit 31-(e%32)."""
    assert bits01.shape[-1] % 32 == 0
    b = np.packbits(bits01.astype(np.uint8), axis=-1, bitorder="big")  # bytes, MSB first
    b = b.reshape(b.shape[:-1] + (b.shape[-1] // 4, 4)).astype(np.uint32)
    return (b[..., 0] << 24) | (b[..., 1] << 16) | (b[..., 2] << 8) | b[..., 3]


def _unpack_msb_first(words: np.ndarray) -> np.ndarray:
    """[..., n] uint32 -> [..., 32*n] uint8 0/1."""
    w = np.asarray(words, dtype=np.uint32)
    by = np.stack([(w >> 24) & 0xFF, (w >> 16) & 0xFF, (w >> 8) & 0xFF, w & 0xFF], axis=-1).astype(np.uint8)
    bits = np.unpackbits(by, axis=-1, bitorder="big")
    return bits.reshape(w.shape[:-1] + (w.shape[-1] * 32,))


def _gather_zero_ext(words: np.ndarray, idx: np.ndarray) -> np.ndarray:
    words = np.asarray(words).reshape(-1).view(np.uint32) if np.asarray(words).dtype != np.uint32 else np.asarray(words).reshape(-1)
    ok = (idx >= 0) & (idx < words.size)
    out = np.zeros(idx.shape, dtype=np.uint32)
    out[ok] = words[idx[ok]]
    return out


# --------------------------------------------------------------------------------------
# val2bit: QGTC_device.cu:44-130
# --------------------------------------------------------------------------------------


def pack_planes(q: np.ndarray, nbits: int, layout: int) -> np.ndarray:
    """Pack planes 0..nbits-1 of the int32 matrix q[H,W] into flat uint32 words."""
    q = np.asarray(q, dtype=np.int64)
    H, W = q.shape
    out = np.zeros(packed_words(layout, nbits, H, W), dtype=np.uint32)
    if layout == LAYOUT_ROW:
        # kernel.h:204-242: word[p][r][w] <- columns 32w.. of row r; rows>=H, cols>=W are 0
        rows, wpr = P8(H), S128(W) * 4
        for p in range(nbits):
            bits = np.zeros((rows, wpr * 32), dtype=np.uint8)
            bits[:H, :W] = (q >> p) & 1
            out[p * rows * wpr:(p + 1) * rows * wpr] = _pack_msb_first(bits).reshape(-1)
        return out
    # kernel.h:75-106: memory [p][col][S128(H)*4]; plane stride ALWAYS uses P128(W) (kernel.h:88)
    kw = S128(H) * 4
    plane_stride = S128(H) * P128(W) * 4
    for p in range(nbits):
        bits = np.zeros((W, kw * 32), dtype=np.uint8)
        bits[:, :H] = ((q >> p) & 1).T
        words = _pack_msb_first(bits)                       # [W, kw]
        idx = p * plane_stride + np.arange(W)[:, None] * kw + np.arange(kw)[None, :]
        ok = idx < out.size                                   # COL8 with P8(W) != P128(W): dropped (rule 10)
        out[idx[ok]] = words[ok]
    return out


def val2bit(x: np.ndarray, nbits: int, col_major: bool, output_layer: bool) -> np.ndarray:
    """Returns the packed tensor with the reference's reported 2-D shape (uint32)."""
    x = np.asarray(x, dtype=np.float32)
    layout = LAYOUT_ROW if not col_major else (LAYOUT_COL8 if output_layer else LAYOUT_COL128)
    flat = pack_planes(quantize_val(x, nbits), nbits, layout)
    return flat.reshape(packed_shape(layout, nbits, x.shape[0], x.shape[1]))


# --------------------------------------------------------------------------------------
# bit2val: QGTC_device.cu:135-206, kernel.h:109-201
# --------------------------------------------------------------------------------------


def bit2val(words: np.ndarray, nbits: int, H: int, W: int, col_major: bool, output_layer: bool = False) -> np.ndarray:
    """Inverse of pack -> int32 [H,W].  Both col-major variants read with the C128 formula
    (UnPackFcWeight128, kernel.h:122,131), exactly as the reference launcher does."""
    flat = np.asarray(words).reshape(-1).view(np.uint32)
    out = np.zeros((H, W), dtype=np.int64)
    if not col_major:
        rows, wpr = P8(H), S128(W) * 4
        for p in range(nbits):
            idx = p * rows * wpr + np.arange(H)[:, None] * wpr + np.arange(wpr)[None, :]
            bits = _unpack_msb_first(_gather_zero_ext(flat, idx))[:, :W]
            out += bits.astype(np.int64) << p
    else:
        kw = S128(H) * 4
        plane_stride = S128(H) * P128(W) * 4
        for p in range(nbits):
            idx = p * plane_stride + np.arange(W)[:, None] * kw + np.arange(kw)[None, :]
            bits = _unpack_msb_first(_gather_zero_ext(flat, idx))[:, :H]   # [W, H]
            out += bits.T.astype(np.int64) << p
    return out.astype(np.int32)


# --------------------------------------------------------------------------------------
# bit-GEMM core: kernel.h:265-342 (shared by all GEMM kernels)
# --------------------------------------------------------------------------------------


def _operand_planes(x1: np.ndarray, x2: np.ndarray, M: int, K: int, N: int, bit1: int, bit2: int, pad128_b: bool):
    """Fetch A planes [bit1, M, Kp] and B planes [bit2, N, Kp] as 0/1 uint8 from raw words."""
    f1 = np.asarray(x1).reshape(-1).view(np.uint32)
    f2 = np.asarray(x2).reshape(-1).view(np.uint32)
    gdk4 = S128(K) * 4
    act_offset = P8(M) * gdk4                               # kernel.h:265
    w_offset = gdk4 * (P128(N) if pad128_b else P8(N))       # kernel.h:266 / :836
    a = np.zeros((bit1, M, gdk4 * 32), dtype=np.uint8)
    b = np.zeros((bit2, N, gdk4 * 32), dtype=np.uint8)
    for p in range(bit1):
        idx = p * act_offset + np.arange(M)[:, None] * gdk4 + np.arange(gdk4)[None, :]
        a[p] = _unpack_msb_first(_gather_zero_ext(f1, idx))
    for q in range(bit2):
        idx = q * w_offset + np.arange(N)[:, None] * gdk4 + np.arange(gdk4)[None, :]
        b[q] = _unpack_msb_first(_gather_zero_ext(f2, idx))
    return a, b


def bitgemm_acc(x1, x2, M: int, K: int, N: int, bit1: int, bit2: int, pad128_b: bool = True) -> np.ndarray:
    """acc[m][n] = sum_{p,q} popc(X1[p][m] & X2[q][n]) << (p+q), wrapping int32 (kernel.h:292-342).

    Only rows m<M and columns n<N are returned: every consumer masks the rest
    (kernel.h:367-368, :921-923)."""
    a, b = _operand_planes(x1, x2, M, K, N, bit1, bit2, pad128_b)
    acc = np.zeros((M, N), dtype=np.int64)
    for q in range(bit2):
        bq = b[q].astype(np.float64).T                      # [Kp, N]
        for p in range(bit1):
            # AND+popc of 0/1 vectors == dot product; exact in float64 (<= Kp < 2^53)
            cnt = (a[p].astype(np.float64) @ bq).astype(np.int64)
            acc += cnt << (p + q)
    return (acc & 0xFFFFFFFF).astype(np.uint32).view(np.int32).reshape(M, N)


def requantize(acc: np.ndarray, out_bit: int) -> np.ndarray:
    """quantize(c, ob, 2^ob, 0): kernel.h:31-37,350.  acc>2^ob -> 2^ob-1; acc<0 -> 1; else acc.
    Only bits 0..ob-1 are packed afterwards, so acc==2^ob contributes 0 bits."""
    acc = np.asarray(acc, dtype=np.int64)
    mx = 1 << out_bit
    return np.where(acc > mx, mx - 1, np.where(acc < 0, 1, acc)).astype(np.int32)


def bitMM2Int(x1, x2, M, K, N, bit1, bit2, pad_128: bool = False) -> np.ndarray:
    """QGTC_layer_output_PAD8 / _PAD128 (kernel.h:816-1054): float(acc) [M,N]."""
    return bitgemm_acc(x1, x2, M, K, N, bit1, bit2, pad128_b=bool(pad_128)).astype(np.float32)


def bitMM2Bit(x1, x2, M, K, N, bit1, bit2, out_bit) -> np.ndarray:
    """QGTC_layer_hidden (kernel.h:245-391): R(out_bit, M, N), shape [out_bit*P8(M), S128(N)*4]."""
    v = requantize(bitgemm_acc(x1, x2, M, K, N, bit1, bit2, True), out_bit)
    return pack_planes(v, out_bit, LAYOUT_ROW).reshape(packed_shape(LAYOUT_ROW, out_bit, M, N))


def bitMM2Bit_col(x1, x2, M, K, N, bit1, bit2, out_bit, ref_quirk: bool = False) -> np.ndarray:
    """QGTC_layer_hidden_col (kernel.h:651-810): output C128(out_bit, M, N), reported shape
    [out_bit*S128(M)*4, P128(N)] (QGTC_device.cu:456).

    ref_quirk=False: the intended semantics -- a correct column-major pack.
    ref_quirk=True : byte-for-byte what the shipped kernel stores, including its two defects
                     (second half read at +4 instead of +32, kernel.h:782; transposed bounds
                     test, kernel.h:777-778)."""
    v = requantize(bitgemm_acc(x1, x2, M, K, N, bit1, bit2, True), out_bit)
    shape = packed_shape(LAYOUT_COL128, out_bit, M, N)
    if not ref_quirk:
        return pack_planes(v, out_bit, LAYOUT_COL128).reshape(shape)
    gdh = S128(M)
    opt_offset_bytes = gdh * P128(N) * 16
    out = np.zeros(shape[0] * shape[1] * 4, dtype=np.uint8)      # byte view, little-endian words
    vp = np.zeros((P8(M), P8(N)), dtype=np.int64)
    vp[:M, :N] = v
    # NB: the kernel quantizes the full 8x8 fragment; rows>=M / cols>=N of acc come from padded
    # operands.  They are only reachable through the defective +4 read; we take them as computed
    # from zero-extended operands (== 0 contribution) which is what well-formed inputs give.
    full = requantize(bitgemm_acc_padded(x1, x2, M, K, N, bit1, bit2), out_bit).astype(np.int64)
    for bx in range(S8(M)):
        for by in range(S8(N)):
            tile = full[bx * 8:bx * 8 + 8, by * 8:by * 8 + 8]
            cs = tile.T.reshape(-1)                                 # col-major store: Cs[col*8+row]
            for p in range(out_bit):
                for half, base in ((0, 0), (1, 4)):
                    for j in range(4):
                        byte = 0
                        for i in range(8):
                            lane = 8 * j + i
                            gy, gx = lane % 8, lane // 8
                            inb = (by * 8 + gy < N) and (bx * 8 + gx + 4 * half < M)
                            bit = (int(cs[base + lane]) >> p) & 1 if inb else 0
                            byte |= bit << (7 - i)
                        col = by * 8 + 4 * half + j
                        out[p * opt_offset_bytes + col * gdh * 16 + (bx ^ 3)] = byte
    return out.view(np.uint32).reshape(shape)


def bitgemm_acc_padded(x1, x2, M, K, N, bit1, bit2) -> np.ndarray:
    """acc over the padded P8(M) x P8(N) fragment grid (what the kernels actually compute)."""
    return bitgemm_acc(x1, x2, P8(M), K, P8(N), bit1, bit2, True) if (P8(M) != M or P8(N) != N) \
        else bitgemm_acc(x1, x2, M, K, N, bit1, bit2, True)


# --------------------------------------------------------------------------------------
# tile counters: kernel.h:394-648 (SURVEY 8c rule 9)
# --------------------------------------------------------------------------------------


def tile_counts(x1, M: int, K: int, N: int, bit1: int, bit2: int) -> tuple[int, int]:
    """(counter_global increment, counter increment) for one call.

    counter_global (kernel.h:452): every lane of every warp adds 1 per (tile, plane pair, k step).
    counter (kernel.h:587): same, but only for non-zero 8x128-bit A tiles.  The shipped zero test
    probes rows with a 32x too large stride (kernel.h:581, ``laneid*gdk*128`` on an int*); we count
    the INTENDED tiles (all 1024 bits of the 8-row x 128-bit tile), see DESIGN.md."""
    f1 = np.asarray(x1).reshape(-1).view(np.uint32)
    gdk = S128(K)
    total = 32 * S8(M) * S8(N) * bit1 * bit2 * gdk
    nz = 0
    act_offset = P8(M) * gdk * 4
    for p in range(bit1):
        idx = p * act_offset + np.arange(P8(M))[:, None] * gdk * 4 + np.arange(gdk * 4)[None, :]
        w = _gather_zero_ext(f1, idx).reshape(S8(M), 8, gdk, 4)
        nz += int((w != 0).any(axis=(1, 3)).sum())
    return total, 32 * S8(N) * bit2 * nz


def tile_bitmap(x1, M: int, K: int) -> np.ndarray:
    """Checker for qgtc_tile_bitmap (this repo's extension of the zero-tile test, kernel.h:574-591, to the
    128-row x 128-bit tile the tensor core consumes): uint32 [ceil(P8(M)/128)][ceil(S128(K)/32)], bit c of word w
    set iff rows 128mt..128mt+127 x chunk 32w+c of plane 0 contain a one."""
    f1 = np.asarray(x1).reshape(-1).view(np.uint32)
    gdk = S128(K)
    mt, bmw = (P8(M) + 127) // 128, (gdk + 31) // 32
    idx = np.arange(P8(M))[:, None] * gdk * 4 + np.arange(gdk * 4)[None, :]
    w = _gather_zero_ext(f1, idx).reshape(P8(M), gdk, 4)
    nzr = (w != 0).any(axis=2)                                   # [rows, chunks]
    pad = np.zeros((mt * 128, bmw * 32), dtype=bool)
    pad[:P8(M), :gdk] = nzr
    tiles = pad.reshape(mt, 128, bmw, 32).any(axis=1)            # [mt, bmw, 32]
    return (tiles.astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, None, :]).sum(axis=2).astype(np.uint32)


# --------------------------------------------------------------------------------------
# value-level identity (property used by tests; not a reference function)
# --------------------------------------------------------------------------------------


def int_gemm_identity(a_val: np.ndarray, b_val: np.ndarray) -> np.ndarray:
    """(A_q @ X_q) on the low-b-bit values, wrapping int32."""
    r = a_val.astype(np.int64) @ b_val.astype(np.int64)
    return (r & 0xFFFFFFFF).astype(np.uint32).view(np.int32)

"""ctypes mirror of the C ABI (include/qgtc_b200.h) on torch CUDA tensors.

Function names, argument order and returned shapes follow the reference's Python surface
(QGTC_module/QGTC_host.cpp:106-256); every call goes through ``libqgtc_b200.so`` -- the same entry
points a cgo/JNI/ctypes binding of another host would use.  No CPU fallback exists: loading fails
loudly when the library is not built.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import lib_dir

ENGINE_AUTO, ENGINE_MMA_SYNC, ENGINE_TCGEN05, ENGINE_GROUPED = 0, 1, 2, 3
GROUP_BITS_ROW, GROUP_FLOAT, GROUP_BITS_COL = 0, 1, 2
FLAG_NO_ZERO_JUMP, FLAG_COL_REF_QUIRK, FLAG_TC05_PAIR, FLAG_X1_STATIC, FLAG_X2_STATIC = 0x1, 0x2, 0x4, 0x8, 0x10
LAYOUT_ROW, LAYOUT_COL128, LAYOUT_COL8 = 0, 1, 2

# every symbol include/qgtc_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "qgtc_abi_version", "qgtc_strerror", "qgtc_last_cuda_error", "qgtc_packed_words", "qgtc_packed_shape",
    "qgtc_val2bit", "qgtc_bit2val", "qgtc_bitmm2bit", "qgtc_bitmm2bit_col", "qgtc_bitmm2int",
    "qgtc_tile_counts", "qgtc_bitmm2bit_batched", "qgtc_bitmm2bit_host", "qgtc_launch_count",
    "qgtc_tile_bitmap_words", "qgtc_tile_bitmap", "qgtc_bitmm2bit_tilemap", "qgtc_bitmm2int_tilemap",
    "qgtc_bitmm2bit_host_batched", "qgtc_edges2bit", "qgtc_bitmm_grouped", "qgtc_bitmm2bit_host_batched_multi",
    "qgtc_bitmm_grouped_fused", "qgtc_val2bit_batched",
]


class HostVariant(ctypes.Structure):
    """qgtc_host_variant (include/qgtc_b200.h)."""
    _fields_ = [("x2_host", ctypes.c_void_p), ("x2_stride_words", ctypes.c_size_t), ("bit2", ctypes.c_int), ("out_bit", ctypes.c_int),
                ("out_host", ctypes.c_void_p), ("out_stride_words", ctypes.c_size_t)]


class FusedUpdate(ctypes.Structure):
    """qgtc_fused_update (include/qgtc_b200.h)."""
    _fields_ = [("w_dev", ctypes.c_void_p), ("w_words", ctypes.c_size_t), ("N2", ctypes.c_int), ("bit_w", ctypes.c_int),
                ("out_bit2", ctypes.c_int), ("kind2", ctypes.c_int), ("pad128", ctypes.c_int)]


class PackProblem(ctypes.Structure):
    """qgtc_pack_problem (include/qgtc_b200.h)."""
    _fields_ = [("in_dev", ctypes.c_void_p), ("out_dev", ctypes.c_void_p), ("out_words", ctypes.c_size_t), ("H", ctypes.c_int), ("W", ctypes.c_int)]


class GroupProblem(ctypes.Structure):
    """qgtc_group_problem (include/qgtc_b200.h)."""
    _fields_ = [("x1_dev", ctypes.c_void_p), ("x1_words", ctypes.c_size_t), ("x2_dev", ctypes.c_void_p), ("x2_words", ctypes.c_size_t),
                ("out_dev", ctypes.c_void_p), ("out_size", ctypes.c_size_t), ("tile_bitmap_dev", ctypes.c_void_p),
                ("M", ctypes.c_int32), ("K", ctypes.c_int32)]

_LIB = None


class QgtcError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(lib_dir(), "libqgtc_b200.so")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError("libqgtc_b200.so is not built (%s); run __graft_entry__.build(). No CPU fallback." % path)
        L = ctypes.CDLL(path)
        c = ctypes
        vp, sz, i = c.c_void_p, c.c_size_t, c.c_int
        L.qgtc_abi_version.restype = i
        L.qgtc_strerror.restype = c.c_char_p
        L.qgtc_strerror.argtypes = [i]
        L.qgtc_last_cuda_error.restype = c.c_char_p
        L.qgtc_packed_words.restype = sz
        L.qgtc_packed_words.argtypes = [i, i, i, i]
        L.qgtc_packed_shape.restype = None
        L.qgtc_packed_shape.argtypes = [i, i, i, i, c.POINTER(c.c_int64), c.POINTER(c.c_int64)]
        L.qgtc_val2bit.argtypes = [vp, i, i, i, i, i, vp, sz, vp]
        L.qgtc_bit2val.argtypes = [vp, sz, i, i, i, i, i, vp, vp]
        L.qgtc_val2bit_batched.argtypes = [vp, i, i, i, i, vp]
        L.qgtc_bitmm2bit.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_bitmm2bit_col.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_bitmm2int.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_tile_counts.argtypes = [vp, sz, i, i, i, i, i, c.POINTER(c.c_uint64), c.POINTER(c.c_uint64), vp]
        L.qgtc_bitmm2bit_batched.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_bitmm2bit_host.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_launch_count.restype = c.c_uint64
        L.qgtc_bitmm2bit_host_batched.argtypes = [vp, sz, vp, sz, i, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_edges2bit.argtypes = [vp, vp, sz, i, i, vp, sz, vp]
        L.qgtc_tile_bitmap_words.restype = sz
        L.qgtc_tile_bitmap_words.argtypes = [i, i, i]
        L.qgtc_tile_bitmap.argtypes = [vp, sz, sz, i, i, i, vp, sz, vp]
        L.qgtc_bitmm2bit_tilemap.argtypes = [vp, sz, vp, sz, vp, i, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_bitmm2int_tilemap.argtypes = [vp, sz, vp, sz, vp, i, i, i, i, i, vp, sz, i, c.c_uint, vp]
        L.qgtc_bitmm2bit_host_batched_multi.argtypes = [vp, sz, i, i, i, i, i, c.POINTER(HostVariant), i, i, c.c_uint, vp]
        L.qgtc_bitmm_grouped_fused.argtypes = [c.POINTER(GroupProblem), i, i, i, i, i, c.POINTER(FusedUpdate), c.POINTER(c.c_void_p),
                                               c.POINTER(c.c_size_t), i, i, c.c_uint, vp]
        L.qgtc_bitmm_grouped.argtypes = [c.POINTER(GroupProblem), i, i, i, i, i, i, i, i, c.c_uint, vp]
        _LIB = L
    return _LIB


def _check(rc: int, op: str) -> None:
    if rc != 0:
        L = lib()
        msg = L.qgtc_strerror(rc).decode()
        if rc == 5:
            msg += ": " + L.qgtc_last_cuda_error().decode()
        raise QgtcError("%s: %s" % (op, msg))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def packed_shape(layout: int, bits: int, H: int, W: int) -> tuple[int, int]:
    r, c = ctypes.c_int64(0), ctypes.c_int64(0)
    lib().qgtc_packed_shape(layout, bits, H, W, ctypes.byref(r), ctypes.byref(c))
    return int(r.value), int(c.value)


def packed_words(layout: int, bits: int, H: int, W: int) -> int:
    return int(lib().qgtc_packed_words(layout, bits, H, W))


def launch_count() -> int:
    return int(lib().qgtc_launch_count())


def _req_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise QgtcError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise QgtcError("%s must be contiguous" % name)
    if t.dtype != dtype:
        raise QgtcError("%s must be %s" % (name, dtype))
    return t


def val2bit(x: torch.Tensor, nbits: int, col_major: bool, output_layer: bool) -> torch.Tensor:
    """QGTC.val2bit (QGTC_host.cpp:229-238)."""
    _req_cuda(x, torch.float32, "input")
    H, W = x.shape
    layout = LAYOUT_ROW if not col_major else (LAYOUT_COL8 if output_layer else LAYOUT_COL128)
    with torch.cuda.device(x.device):
        out = torch.empty(packed_shape(layout, nbits, H, W), dtype=torch.int32, device=x.device)
        _check(lib().qgtc_val2bit(x.data_ptr(), H, W, nbits, int(col_major), int(output_layer), out.data_ptr(), out.numel(), _stream()), "val2bit")
    return out


def val2bit_batched(xs, nbits: int, col_major: bool, output_layer: bool):
    """qgtc_val2bit_batched: val2bit of a list of fp32 matrices in one call (ROW layout: one launch for the whole list).
    Returns the list of packed tensors QGTC.val2bit would return for each of them."""
    if len(xs) == 0:
        return []
    for x in xs:
        _req_cuda(x, torch.float32, "input")
    layout = LAYOUT_ROW if not col_major else (LAYOUT_COL8 if output_layer else LAYOUT_COL128)
    dev = xs[0].device
    with torch.cuda.device(dev):
        outs = [torch.empty(packed_shape(layout, nbits, x.shape[0], x.shape[1]), dtype=torch.int32, device=dev) for x in xs]
        arr = (PackProblem * len(xs))()
        for i, (x, o) in enumerate(zip(xs, outs)):
            arr[i].in_dev, arr[i].out_dev, arr[i].out_words = x.data_ptr(), o.data_ptr(), o.numel()
            arr[i].H, arr[i].W = x.shape[0], x.shape[1]
        _check(lib().qgtc_val2bit_batched(ctypes.addressof(arr), len(xs), nbits, int(col_major), int(output_layer), _stream()), "val2bit_batched")
    return outs


def bit2val(w: torch.Tensor, nbits: int, H: int, W: int, col_major: bool, output_layer: bool = False) -> torch.Tensor:
    """QGTC.bit2val (QGTC_host.cpp:244-256)."""
    _req_cuda(w, torch.int32, "input")
    with torch.cuda.device(w.device):
        out = torch.empty((H, W), dtype=torch.int32, device=w.device)
        _check(lib().qgtc_bit2val(w.data_ptr(), w.numel(), nbits, H, W, int(col_major), int(output_layer), out.data_ptr(), _stream()), "bit2val")
    return out


def bitMM2Bit(x1, x2, M, K, N, bit1, bit2, out_bit, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """QGTC.bitMM2Bit (QGTC_host.cpp:106-122)."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2")
    with torch.cuda.device(x1.device):
        out = torch.empty(packed_shape(LAYOUT_ROW, out_bit, M, N), dtype=torch.int32, device=x1.device)
        _check(lib().qgtc_bitmm2bit(x1.data_ptr(), x1.numel(), x2.data_ptr(), x2.numel(), M, K, N, bit1, bit2, out_bit,
                                    out.data_ptr(), out.numel(), engine, flags, _stream()), "bitMM2Bit")
    return out


def bitMM2Bit_col(x1, x2, M, K, N, bit1, bit2, out_bit, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """QGTC.bitMM2Bit_col (QGTC_host.cpp:185-201); flags=FLAG_COL_REF_QUIRK for the shipped kernel's bytes."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2")
    with torch.cuda.device(x1.device):
        out = torch.empty(packed_shape(LAYOUT_COL128, out_bit, M, N), dtype=torch.int32, device=x1.device)
        _check(lib().qgtc_bitmm2bit_col(x1.data_ptr(), x1.numel(), x2.data_ptr(), x2.numel(), M, K, N, bit1, bit2, out_bit,
                                        out.data_ptr(), out.numel(), engine, flags, _stream()), "bitMM2Bit_col")
    return out


def bitMM2Int(x1, x2, M, K, N, bit1, bit2, pad_128=False, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """QGTC.bitMM2Int (QGTC_host.cpp:207-223)."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2")
    with torch.cuda.device(x1.device):
        out = torch.empty((M, N), dtype=torch.float32, device=x1.device)
        _check(lib().qgtc_bitmm2int(x1.data_ptr(), x1.numel(), x2.data_ptr(), x2.numel(), M, K, N, bit1, bit2, int(bool(pad_128)),
                                    out.data_ptr(), out.numel(), engine, flags, _stream()), "bitMM2Int")
    return out


def tile_counts(x1, M, K, N, bit1, bit2) -> tuple[int, int]:
    """(counter_global, counter) increments of bitMM2Bit_base_cnt / _zerojump_cnt (QGTC_device.cu:272-373)."""
    _req_cuda(x1, torch.int32, "bit_X1")
    t, z = ctypes.c_uint64(0), ctypes.c_uint64(0)
    with torch.cuda.device(x1.device):
        _check(lib().qgtc_tile_counts(x1.data_ptr(), x1.numel(), M, K, N, bit1, bit2, ctypes.byref(t), ctypes.byref(z), _stream()), "tile_counts")
    return int(t.value), int(z.value)


def bitMM2Bit_batched(x1, x2, count, M, K, N, bit1, bit2, out_bit, out=None, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """`count` independent problems in one launch.  x1: [count, words1], x2: [count, words2] (or a
    single problem's words to alias it for every problem); returns [count, words_out]."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2")
    w1, w2, wo = packed_words(LAYOUT_ROW, bit1, M, K), packed_words(LAYOUT_COL128, bit2, K, N), packed_words(LAYOUT_ROW, out_bit, M, N)
    s1 = 0 if x1.numel() == w1 else w1
    s2 = 0 if x2.numel() == w2 else w2
    if (s1 and x1.numel() != count * w1) or (s2 and x2.numel() != count * w2):
        raise QgtcError("batched operands must hold `count` well-formed problems")
    with torch.cuda.device(x1.device):
        if out is None:
            out = torch.empty((count, wo), dtype=torch.int32, device=x1.device)
        else:
            _req_cuda(out, torch.int32, "out")
            if out.numel() < count * wo:
                raise QgtcError("out must hold count * %d words (got %d)" % (wo, out.numel()))
        _check(lib().qgtc_bitmm2bit_batched(x1.data_ptr(), s1, x2.data_ptr(), s2, count, M, K, N, bit1, bit2, out_bit,
                                            out.data_ptr(), wo, engine, flags, _stream()), "bitMM2Bit_batched")
    return out


def bitMM2Bit_aliased(x1, x2, count, M, K, N, bit1, bit2, out_bit, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """`count` repetitions of ONE problem in one launch, all writing the same output (out stride 0): the throughput
    form of QGTC.bitMM2Bit_profile (QGTC_device.cu:403-418, 200 back-to-back launches on the same operands)."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2")
    with torch.cuda.device(x1.device):
        out = torch.empty(packed_shape(LAYOUT_ROW, out_bit, M, N), dtype=torch.int32, device=x1.device)
        _check(lib().qgtc_bitmm2bit_batched(x1.data_ptr(), 0, x2.data_ptr(), 0, count, M, K, N, bit1, bit2, out_bit,
                                            out.data_ptr(), 0, engine, flags, _stream()), "bitMM2Bit_aliased")
    return out


def tile_bitmap(x1: torch.Tensor, M: int, K: int, count: int = 1) -> torch.Tensor:
    """Tile bitmap of `count` packed 1-bit adjacencies (x1: [count, words] or one problem's words)."""
    _req_cuda(x1, torch.int32, "bit_X1")
    w1 = packed_words(LAYOUT_ROW, 1, M, K)
    stride = 0 if count == 1 else w1
    words = int(lib().qgtc_tile_bitmap_words(M, K, count))
    with torch.cuda.device(x1.device):
        bm = torch.empty((count, words // count), dtype=torch.int32, device=x1.device)
        _check(lib().qgtc_tile_bitmap(x1.data_ptr(), x1.numel() if count == 1 else w1, stride, M, K, count, bm.data_ptr(), words, _stream()), "tile_bitmap")
    return bm


def bitMM2Bit_tilemap(x1, x2, tile_map, count, M, K, N, bit2, out_bit, out=None, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """bitMM2Bit (1-bit X1) for `count` problems with a precomputed tile bitmap: zero tiles are not loaded."""
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2"); _req_cuda(tile_map, torch.int32, "tile_map")
    w1, w2, wo = packed_words(LAYOUT_ROW, 1, M, K), packed_words(LAYOUT_COL128, bit2, K, N), packed_words(LAYOUT_ROW, out_bit, M, N)
    if x1.numel() != count * w1 or x2.numel() not in (w2, count * w2):
        raise QgtcError("tilemap operands must hold `count` well-formed problems")
    s2 = 0 if (count > 1 and x2.numel() == w2) else w2
    if tile_map.numel() < int(lib().qgtc_tile_bitmap_words(M, K, count)):
        raise QgtcError("tile_map holds %d words, the problem needs %d" % (tile_map.numel(), int(lib().qgtc_tile_bitmap_words(M, K, count))))
    with torch.cuda.device(x1.device):
        if out is None:
            out = torch.empty((count, wo), dtype=torch.int32, device=x1.device)
        else:
            _req_cuda(out, torch.int32, "out")
            if out.numel() < count * wo:
                raise QgtcError("out must hold count * %d words (got %d)" % (wo, out.numel()))
        _check(lib().qgtc_bitmm2bit_tilemap(x1.data_ptr(), w1, x2.data_ptr(), s2, tile_map.data_ptr(), count, M, K, N, bit2, out_bit,
                                            out.data_ptr(), wo, engine, flags, _stream()), "bitMM2Bit_tilemap")
    return out


def bitMM2Int_tilemap(x1, x2, tile_map, M, K, N, bit2, pad_128=False, engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    _req_cuda(x1, torch.int32, "bit_X1"); _req_cuda(x2, torch.int32, "bit_X2"); _req_cuda(tile_map, torch.int32, "tile_map")
    if tile_map.numel() < int(lib().qgtc_tile_bitmap_words(M, K, 1)):
        raise QgtcError("tile_map holds %d words, the problem needs %d" % (tile_map.numel(), int(lib().qgtc_tile_bitmap_words(M, K, 1))))
    with torch.cuda.device(x1.device):
        out = torch.empty((M, N), dtype=torch.float32, device=x1.device)
        _check(lib().qgtc_bitmm2int_tilemap(x1.data_ptr(), x1.numel(), x2.data_ptr(), x2.numel(), tile_map.data_ptr(), M, K, N, bit2,
                                            int(bool(pad_128)), out.data_ptr(), out.numel(), engine, flags, _stream()), "bitMM2Int_tilemap")
    return out


def bitMM2Bit_host(x1_host: torch.Tensor, x2_host: torch.Tensor, M, K, N, bit1, bit2, out_bit, out_host=None,
                   engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """Host-buffer entry point (pinned CPU int32 tensors in, CPU tensor out): H2D + GEMM + D2H + sync."""
    for t, n in ((x1_host, "bit_X1"), (x2_host, "bit_X2")):
        if t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
            raise QgtcError("%s must be a contiguous int32 CPU tensor" % n)
    if out_host is None:
        out_host = torch.empty(packed_shape(LAYOUT_ROW, out_bit, M, N), dtype=torch.int32).pin_memory()
    _check(lib().qgtc_bitmm2bit_host(x1_host.data_ptr(), x1_host.numel(), x2_host.data_ptr(), x2_host.numel(), M, K, N,
                                     bit1, bit2, out_bit, out_host.data_ptr(), out_host.numel(), engine, flags, _stream()), "bitMM2Bit_host")
    return out_host


def bitMM2Bit_host_batched(x1_host: torch.Tensor, x2_host: torch.Tensor, count, M, K, N, bit1, bit2, out_bit, out_host=None,
                           engine=ENGINE_AUTO, flags=0) -> torch.Tensor:
    """Pipelined host-buffer entry point: [count, words] pinned CPU tensors in, [count, words] CPU tensor out."""
    for t, n in ((x1_host, "bit_X1"), (x2_host, "bit_X2")):
        if t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous() or t.dim() != 2 or t.shape[0] != count:
            raise QgtcError("%s must be a contiguous [count, words] int32 CPU tensor" % n)
    wo = packed_words(LAYOUT_ROW, out_bit, M, N)
    if out_host is None:
        out_host = torch.empty((count, wo), dtype=torch.int32).pin_memory()
    _check(lib().qgtc_bitmm2bit_host_batched(x1_host.data_ptr(), x1_host.shape[1], x2_host.data_ptr(), x2_host.shape[1], count, M, K, N,
                                             bit1, bit2, out_bit, out_host.data_ptr(), out_host.shape[1], engine, flags, _stream()),
           "bitMM2Bit_host_batched")
    return out_host


def edges2bit(rows: torch.Tensor, cols: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """Packed 1-bit adjacency R(1,H,W) from int32 edge endpoints on the device (== val2bit of the dense 0/1 matrix)."""
    _req_cuda(rows, torch.int32, "rows"); _req_cuda(cols, torch.int32, "cols")
    if rows.numel() != cols.numel():
        raise QgtcError("rows and cols must have the same length")
    with torch.cuda.device(rows.device):
        out = torch.empty(packed_shape(LAYOUT_ROW, 1, H, W), dtype=torch.int32, device=rows.device)
        _check(lib().qgtc_edges2bit(rows.data_ptr(), cols.data_ptr(), rows.numel(), H, W, out.data_ptr(), out.numel(), _stream()), "edges2bit")
    return out


def bitmm_grouped(x1s, x2s, Ms, Ks, N, bit1, bit2, out_bit, kind=GROUP_BITS_ROW, pad_128=False, tile_maps=None, outs=None,
                  engine=ENGINE_AUTO, flags=0):
    """One launch for len(x1s) independent problems sharing N and the bit widths (qgtc_bitmm_grouped): problem i is
    x1s[i] (R(bit1, Ms[i], Ks[i]) raw words) x x2s[i] (C(bit2, Ks[i], N) raw words).  Results equal bitMM2Bit /
    bitMM2Int / bitMM2Bit_col of each problem.  Returns the list of outputs (views of one allocation unless `outs` is given)."""
    count = len(x1s)
    if not (count == len(x2s) == len(Ms) == len(Ks)) or count == 0:
        raise QgtcError("bitmm_grouped: operand lists must be non-empty and of equal length")
    dev = x1s[0].device
    arr = (GroupProblem * count)()
    if kind == GROUP_FLOAT:
        sizes = [int(m) * N for m in Ms]
        shapes = [(int(m), N) for m in Ms]
        dtype = torch.float32
    else:
        layout = LAYOUT_ROW if kind == GROUP_BITS_ROW else LAYOUT_COL128
        shapes = [packed_shape(layout, out_bit, int(m), N) for m in Ms]
        sizes = [a * b for a, b in shapes]
        dtype = torch.int32
    with torch.cuda.device(dev):
        if outs is None:
            offs, total = [], 0
            for sz_ in sizes:
                offs.append(total)
                total += (sz_ + 3) & ~3                 # keep every problem's output 16-byte aligned
            flat = torch.empty((total,), dtype=dtype, device=dev)
            outs = [flat[o:o + sz_].view(shp) for o, sz_, shp in zip(offs, sizes, shapes)]
        for k in range(count):
            a, b, o = x1s[k], x2s[k], outs[k]
            _req_cuda(a, torch.int32, "bit_X1[%d]" % k); _req_cuda(b, torch.int32, "bit_X2[%d]" % k); _req_cuda(o, dtype, "out[%d]" % k)
            p = arr[k]
            p.x1_dev, p.x1_words, p.x2_dev, p.x2_words = a.data_ptr(), a.numel(), b.data_ptr(), b.numel()
            p.out_dev, p.out_size = o.data_ptr(), o.numel()
            p.tile_bitmap_dev = tile_maps[k].data_ptr() if tile_maps is not None and tile_maps[k] is not None else None
            p.M, p.K = int(Ms[k]), int(Ks[k])
        _check(lib().qgtc_bitmm_grouped(arr, count, N, bit1, bit2, out_bit, kind, int(bool(pad_128)), engine, flags, _stream()), "bitmm_grouped")
    return outs


def bitMM2Bit_host_batched_multi(x1_host, x2_hosts, count, M, K, N, bit1, bit2s, out_bits, out_hosts=None, engine=ENGINE_AUTO, flags=0):
    """One pipelined pass over `count` graphs for several feature matrices per graph (qgtc_bitmm2bit_host_batched_multi): the
    adjacency is uploaded once per graph.  x1_host: [count, words]; x2_hosts[v]: [count, words_v]; returns the list of outputs."""
    nvar = len(x2_hosts)
    if nvar == 0 or nvar != len(bit2s) or nvar != len(out_bits):
        raise QgtcError("variant lists must be non-empty and of equal length")
    for t, n in [(x1_host, "bit_X1")] + [(x, "bit_X2[%d]" % v) for v, x in enumerate(x2_hosts)]:
        if t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous() or t.dim() != 2 or t.shape[0] != count:
            raise QgtcError("%s must be a contiguous [count, words] int32 CPU tensor" % n)
    if out_hosts is None:
        out_hosts = [torch.empty((count, packed_words(LAYOUT_ROW, ob, M, N)), dtype=torch.int32).pin_memory() for ob in out_bits]
    arr = (HostVariant * nvar)()
    for v in range(nvar):
        arr[v].x2_host, arr[v].x2_stride_words = x2_hosts[v].data_ptr(), x2_hosts[v].shape[1]
        arr[v].bit2, arr[v].out_bit = int(bit2s[v]), int(out_bits[v])
        arr[v].out_host, arr[v].out_stride_words = out_hosts[v].data_ptr(), out_hosts[v].shape[1]
    _check(lib().qgtc_bitmm2bit_host_batched_multi(x1_host.data_ptr(), x1_host.shape[1], count, M, K, N, bit1, arr, nvar, engine, flags, _stream()),
           "bitMM2Bit_host_batched_multi")
    return out_hosts


def bitmm_grouped_fused(x1s, x2s, Ms, Ks, N, bit1, bit2, out_bit, W, N2, bit_w, out_bit2, kind2=GROUP_BITS_ROW, pad_128=False,
                        tile_maps=None, write_out1=False, engine=ENGINE_AUTO, flags=0):
    """Fused aggregation + update in one launch (qgtc_bitmm_grouped_fused): for every problem
    out2 = op2(bitMM2Bit(x1, x2, M, K, N, bit1, bit2, out_bit), W) with op2 = bitMM2Bit(o, W, M, N, N2, out_bit, bit_w, out_bit2)
    (kind2 = GROUP_BITS_ROW) or bitMM2Int(o, W, M, N, N2, out_bit, bit_w, pad_128) (GROUP_FLOAT).  Returns (outs2, outs1 or None)."""
    count = len(x1s)
    if not (count == len(x2s) == len(Ms) == len(Ks)) or count == 0:
        raise QgtcError("bitmm_grouped_fused: operand lists must be non-empty and of equal length")
    _req_cuda(W, torch.int32, "bit_W")
    dev = x1s[0].device
    arr = (GroupProblem * count)()
    out2_ptrs, out2_sizes = (ctypes.c_void_p * count)(), (ctypes.c_size_t * count)()
    if kind2 == GROUP_FLOAT:
        shapes2, dtype2 = [(int(m), N2) for m in Ms], torch.float32
    else:
        shapes2, dtype2 = [packed_shape(LAYOUT_ROW, out_bit2, int(m), N2) for m in Ms], torch.int32
    shapes1 = [packed_shape(LAYOUT_ROW, out_bit, int(m), N) for m in Ms]

    def alloc(shapes, dtype):
        offs, total = [], 0
        for a, b in shapes:
            offs.append(total)
            total += (a * b + 3) & ~3
        flat = torch.empty((total,), dtype=dtype, device=dev)
        return [flat[o:o + a * b].view((a, b)) for o, (a, b) in zip(offs, shapes)]

    with torch.cuda.device(dev):
        outs2 = alloc(shapes2, dtype2)
        outs1 = alloc(shapes1, torch.int32) if write_out1 else None
        for k in range(count):
            a, b = x1s[k], x2s[k]
            _req_cuda(a, torch.int32, "bit_X1[%d]" % k); _req_cuda(b, torch.int32, "bit_X2[%d]" % k)
            p = arr[k]
            p.x1_dev, p.x1_words, p.x2_dev, p.x2_words = a.data_ptr(), a.numel(), b.data_ptr(), b.numel()
            p.out_dev, p.out_size = (outs1[k].data_ptr(), outs1[k].numel()) if write_out1 else (None, 0)
            p.tile_bitmap_dev = tile_maps[k].data_ptr() if tile_maps is not None and tile_maps[k] is not None else None
            p.M, p.K = int(Ms[k]), int(Ks[k])
            out2_ptrs[k], out2_sizes[k] = outs2[k].data_ptr(), outs2[k].numel()
        upd = FusedUpdate(W.data_ptr(), W.numel(), N2, bit_w, out_bit2, kind2, int(bool(pad_128)))
        _check(lib().qgtc_bitmm_grouped_fused(arr, count, N, bit1, bit2, out_bit, ctypes.byref(upd), out2_ptrs, out2_sizes, int(bool(write_out1)),
                                              engine, flags, _stream()), "bitmm_grouped_fused")
    return outs2, outs1

"""ctypes front-end of oracle/oracle.c (TEST INFRASTRUCTURE ONLY; see oracle.c header).

Same function names and argument meaning as bitplane_oracle.py; used where numpy is too slow
(full-size parity) and as bench.py's cpu_baseline ("port")."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LAYOUT_ROW, LAYOUT_COL128, LAYOUT_COL8 = 0, 1, 2


def build() -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        c = ctypes
        u32p, f32p, i32p = c.POINTER(c.c_uint32), c.POINTER(c.c_float), c.POINTER(c.c_int32)
        L.oracle_num_threads.restype = c.c_int
        L.oracle_packed_words.restype = c.c_size_t
        L.oracle_packed_words.argtypes = [c.c_int] * 4
        L.oracle_val2bit.argtypes = [f32p, c.c_int, c.c_int, c.c_int, c.c_int, u32p, c.c_size_t]
        L.oracle_bit2val.argtypes = [u32p, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_int, i32p]
        L.oracle_bitgemm_acc.argtypes = [u32p, c.c_size_t, u32p, c.c_size_t] + [c.c_int] * 8 + [i32p]
        L.oracle_bitmm2bit.argtypes = [u32p, c.c_size_t, u32p, c.c_size_t] + [c.c_int] * 6 + [u32p, c.c_size_t]
        L.oracle_bitmm2int.argtypes = [u32p, c.c_size_t, u32p, c.c_size_t] + [c.c_int] * 6 + [f32p]
        L.oracle_bitmm2bit_col.argtypes = [u32p, c.c_size_t, u32p, c.c_size_t] + [c.c_int] * 7 + [u32p, c.c_size_t]
        L.oracle_tile_counts.argtypes = [u32p, c.c_size_t] + [c.c_int] * 5 + [c.POINTER(c.c_uint64)] * 2
        L.oracle_tile_counts.restype = None
        _LIB = L
    return _LIB


def _u32(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint32:
        a = a.view(np.uint32) if a.dtype == np.int32 else a.astype(np.uint32)
    return a.reshape(-1)


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def S8(x): return (x + 7) >> 3
def S128(x): return (x + 127) >> 7
def P8(x): return S8(x) << 3
def P128(x): return S128(x) << 7


def packed_shape(layout, bits, H, W):
    if layout == LAYOUT_ROW:
        return (bits * P8(H), S128(W) * 4)
    if layout == LAYOUT_COL128:
        return (bits * S128(H) * 4, P128(W))
    return (bits * S128(H) * 4, P8(W))


def num_threads() -> int:
    return lib().oracle_num_threads()


def val2bit(x, nbits, col_major, output_layer):
    x = np.ascontiguousarray(x, dtype=np.float32)
    H, W = x.shape
    layout = LAYOUT_ROW if not col_major else (LAYOUT_COL8 if output_layer else LAYOUT_COL128)
    shape = packed_shape(layout, nbits, H, W)
    out = np.empty(shape[0] * shape[1], dtype=np.uint32)
    rc = lib().oracle_val2bit(_p(x, ctypes.c_float), H, W, nbits, layout, _p(out, ctypes.c_uint32), out.size)
    assert rc == 0, rc
    return out.reshape(shape)


def bit2val(words, nbits, H, W, col_major, output_layer=False):
    w = _u32(words)
    out = np.empty((H, W), dtype=np.int32)
    lib().oracle_bit2val(_p(w, ctypes.c_uint32), w.size, nbits, H, W, int(bool(col_major)), _p(out, ctypes.c_int32))
    return out


def bitgemm_acc(x1, x2, M, K, N, bit1, bit2, pad128_b=True):
    a, b = _u32(x1), _u32(x2)
    acc = np.empty((M, N), dtype=np.int32)
    rc = lib().oracle_bitgemm_acc(_p(a, ctypes.c_uint32), a.size, _p(b, ctypes.c_uint32), b.size,
                                  M, K, N, bit1, bit2, int(bool(pad128_b)), M, N, _p(acc, ctypes.c_int32))
    assert rc == 0, rc
    return acc


def bitMM2Bit(x1, x2, M, K, N, bit1, bit2, out_bit):
    a, b = _u32(x1), _u32(x2)
    shape = packed_shape(LAYOUT_ROW, out_bit, M, N)
    out = np.empty(shape[0] * shape[1], dtype=np.uint32)
    rc = lib().oracle_bitmm2bit(_p(a, ctypes.c_uint32), a.size, _p(b, ctypes.c_uint32), b.size,
                                M, K, N, bit1, bit2, out_bit, _p(out, ctypes.c_uint32), out.size)
    assert rc == 0, rc
    return out.reshape(shape)


def bitMM2Int(x1, x2, M, K, N, bit1, bit2, pad_128=False):
    a, b = _u32(x1), _u32(x2)
    out = np.empty((M, N), dtype=np.float32)
    rc = lib().oracle_bitmm2int(_p(a, ctypes.c_uint32), a.size, _p(b, ctypes.c_uint32), b.size,
                                M, K, N, bit1, bit2, int(bool(pad_128)), _p(out, ctypes.c_float))
    assert rc == 0, rc
    return out


def bitMM2Bit_col(x1, x2, M, K, N, bit1, bit2, out_bit, ref_quirk=False):
    a, b = _u32(x1), _u32(x2)
    shape = packed_shape(LAYOUT_COL128, out_bit, M, N)
    out = np.empty(shape[0] * shape[1], dtype=np.uint32)
    rc = lib().oracle_bitmm2bit_col(_p(a, ctypes.c_uint32), a.size, _p(b, ctypes.c_uint32), b.size,
                                    M, K, N, bit1, bit2, out_bit, int(bool(ref_quirk)),
                                    _p(out, ctypes.c_uint32), out.size)
    assert rc == 0, rc
    return out.reshape(shape)


def tile_counts(x1, M, K, N, bit1, bit2):
    a = _u32(x1)
    t, z = ctypes.c_uint64(0), ctypes.c_uint64(0)
    lib().oracle_tile_counts(_p(a, ctypes.c_uint32), a.size, M, K, N, bit1, bit2, ctypes.byref(t), ctypes.byref(z))
    return int(t.value), int(z.value)

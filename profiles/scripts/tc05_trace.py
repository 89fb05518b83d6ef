#!/usr/bin/env python3
"""clock64 trace of CTA 0 of the tcgen05 engine (QGTC_TC05_TRACE=1): per pipeline item, where the
producer warp 0, the MMA issuer and epilogue warp 8 spend their cycles."""
import ctypes
import os
import sys

os.environ["QGTC_TC05_TRACE"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from engine_sweep import run  # noqa: E402

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    flags = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
    ms, tops = run(capi.ENGINE_TCGEN05, 4096, 4096, N, 2, 128, reps=1, flags=flags)
    print("ms", ms, "tops", tops)
    buf = np.zeros(3 * 512 * 8, dtype=np.uint64)
    L = capi.lib()
    L.qgtc_debug_tc05_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    n = L.qgtc_debug_tc05_trace(buf.ctypes.data, buf.size)
    t = buf.reshape(3, 512, 8).astype(np.int64)
    t0 = t[1, 0, 0]
    prod, mma, epi = t[0], t[1], t[2]
    print("MMA issuer: item  issue(chunks 0-2)  acquire_next  issue(last)+commit  | loop period")
    for i in range(20, 44):
        r = mma[i]
        print("  %3d  %6d %6d %6d | %6d   (t=%d)" % (i, r[1] - r[0], r[2] - r[1], r[3] - r[2], mma[i + 1][0] - r[0], r[0] - t0))
    print("producer warp 0: item  cpwait+lds  fetch  wait_empty  expand+st  wait_st  arrive | period")
    for i in range(10, 24):
        r = prod[i]
        print("  %3d  %6d %6d %6d %6d %6d %6d | %6d  (t=%d)" % (i, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], prod[i + 1][0] - r[0], r[0] - t0))
    print("epilogue warp 8: tile  wait_dfull  work | period")
    for i in range(0, 6):
        r = epi[i]
        print("  %3d  %7d %7d | %7d" % (i, r[1] - r[0], r[2] - r[1], epi[i + 1][0] - r[0]))
    for name, a, k in (("mma period", mma, 0), ("producer period", prod, 0)):
        d = np.diff(a[8:200, k])
        print(name, "mean", d.mean(), "median", np.median(d))

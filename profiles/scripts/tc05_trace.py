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
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    bits = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    ms, tops = run(capi.ENGINE_TCGEN05, M, M, N, bits, 128 if M >= 4096 else 1024, reps=1, flags=flags)
    print("ms", ms, "tops", tops)
    buf = np.zeros(8 * 512 * 8, dtype=np.uint64)
    L = capi.lib()
    L.qgtc_debug_tc05_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    n = L.qgtc_debug_tc05_trace(buf.ctypes.data, buf.size)
    t = buf.reshape(8, 512, 8).astype(np.int64)
    t0 = t[1, 0, 0]
    prod, mma, epi = t[0], t[1], t[2]
    for role, name in ((1, "MMA issuer 0"), (5, "MMA issuer 1")):
        mm = t[role]
        print(name + " (own items): item  wait_full  flags+ballot  issue+commit | (t_start)")
        for i in range(20, 36):
            r = mm[i]
            if r[0]:
                print("  %3d  %6d %6d %6d |  (t=%d)" % (i, r[4] - r[0], r[1] - r[4], r[3] - r[1], r[0] - t0))
    for role, name in ((1, "MMA issuer 0"), (5, "MMA issuer 1")):
        mm = t[role]
        print(name + " (items of the other issuer): item  t_begin_wait  wait_full")
        for i in range(20, 36):
            r = mm[i]
            if r[5]:
                print("  %3d  %8d %6d" % (i, r[5] - t0, r[6] - r[5]))
    for role, name in ((0, "expander warp 0 (group 0, SMSP0)"), (3, "expander warp 1 (group 0, SMSP1)"), (4, "expander warp 3 (group 0, SMSP3)"), (6, "expander warp 8 (group 1, SMSP0)")):
        pr = t[role]
        print(name + " own items: rawwait+lds  expand0  wait_empty  st+expand1+st  wait_st  flags+arrive | period (t_start, t_arrive)")
        for i in range(10, 20):
            r = pr[i]
            print("  %3d  %6d %6d %6d %6d %6d %6d | %6d  (%d, %d)" % (i, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], pr[i + 1][0] - r[0], r[0] - t0, r[6] - t0))
    print("epilogue warp 8: tile  wait_dfull  work | period")
    for i in range(0, 6):
        r = epi[i]
        print("  %3d  %7d %7d | %7d" % (i, r[1] - r[0], r[2] - r[1], epi[i + 1][0] - r[0]))
    d = np.diff(prod[8:100, 0])
    print("expander own-item period: mean", d.mean(), "median", np.median(d))

#!/usr/bin/env python3
"""clock64 timeline of CTA 0 of the grouped engine (debug instance, QGTC_GRP_TRACE=1) for one op of a ppi-shaped epoch.
Usage: QGTC_GRP_TRACE=1 python profiles/scripts/grp_trace.py [agg|xw] [n] [batches] [bits] [N]"""
import ctypes
import os
import sys

os.environ.setdefault("QGTC_GRP_TRACE", "1")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "agg"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 759
B = int(sys.argv[3]) if len(sys.argv) > 3 else 75
b = int(sys.argv[4]) if len(sys.argv) > 4 else 2
N = int(sys.argv[5]) if len(sys.argv) > 5 else 16
torch.manual_seed(1)
if kind == "agg":
    x1 = [capi.val2bit((torch.rand(n, n, device="cuda") < 0.02).float(), 1, False, False) for _ in range(B)]
    x2 = [capi.val2bit(torch.randint(0, 1 << b, (n, N), device="cuda").float(), b, True, False) for _ in range(B)]
    Ks, b1 = [n] * B, 1
else:
    D = 50
    x1 = [capi.val2bit(torch.randint(0, 1 << b, (n, D), device="cuda").float(), b, False, False) for _ in range(B)]
    w = capi.val2bit(torch.ones(D, N, device="cuda"), b, True, False)
    x2 = [w] * B
    Ks, b1 = [D] * B, b
for _ in range(3):
    outs = capi.bitmm_grouped(x1, x2, [n] * B, Ks, N, b1, b, b, engine=capi.ENGINE_GROUPED)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    capi.bitmm_grouped(x1, x2, [n] * B, Ks, N, b1, b, b, engine=capi.ENGINE_GROUPED)
e1.record()
torch.cuda.synchronize()
print("avg ms per launch (eager, traced instance):", e0.elapsed_time(e1) / 10)
L = capi.lib()
words = 8 * 256 * 2
buf = (ctypes.c_ulonglong * words)()
L.qgtc_debug_grp_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
got = L.qgtc_debug_grp_trace(buf, words)
NAMES = {1: "kernel start", 2: "after setup sync", 3: "after griddep wait", 4: "kernel end", 5: "setup: barriers + tmem alloc issued (warp 1)", 6: "setup: unit records done (warp 1)", 7: "setup: first sync passed",
         10: "exp begin", 11: "exp loads issued", 12: "exp loads landed", 13: "exp chunk0 expanded", 14: "exp stage free", 15: "exp full arrive",
         20: "conv begin", 21: "conv loads issued", 22: "conv loads landed", 23: "conv stage free", 24: "conv full arrive",
         30: "iss unit", 31: "iss acc free", 32: "iss item full", 33: "iss mmas issued",
         40: "epi unit", 41: "epi acc full", 42: "epi done"}
ROLES = {0: "expander w0 (g0,q0)", 1: "expander w4 (g1,q0)", 2: "converter w0", 3: "issuer", 4: "epilogue q0", 7: "cta"}
t0 = None
ev = []
for role in range(8):
    for e in range(256):
        tag, clk = buf[(role * 256 + e) * 2], buf[(role * 256 + e) * 2 + 1]
        if tag == 0:
            break
        ev.append((clk, role, tag))
        if tag == 1:
            t0 = clk
ev.sort()
print("events:", len(ev))
for clk, role, tag in ev:
    print("%8d  %-22s %s" % (clk - t0, ROLES.get(role, role), NAMES.get(tag, tag)))

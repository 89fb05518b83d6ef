#!/usr/bin/env python3
"""Batched val2bit (qgtc_val2bit_batched) against one QGTC.val2bit launch per matrix: the dense batch adjacencies of a Cluster-GCN
epoch (75 matrices of n x n fp32, 1 bit), CUDA events, inputs larger than L2.  Prints GB/s of algorithmic traffic (4 B read +
nbits / 8 B written per element) and the fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

try:
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    HBM = 6650.0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, B in ((759, 75), (1283, 75), (2258, 75)):
    for nbits in (1, 2, 8):
        xs = [(torch.rand(n, n, device="cuda") < 0.02).float() * 3.0 for _ in range(B)]
        by = sum(x.numel() for x in xs) * (4.0 + nbits / 8.0)
        ms_b = timed(lambda: capi.val2bit_batched(xs, nbits, False, False))
        ms_1 = timed(lambda: [capi.val2bit(x, nbits, False, False) for x in xs])
        # the calls above include the output allocations of the Python wrappers: also time graph replays of both forms
        g_b, g_1 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g_b, stream=s):
                keep_b = capi.val2bit_batched(xs, nbits, False, False)
            with torch.cuda.graph(g_1, stream=s):
                keep_1 = [capi.val2bit(x, nbits, False, False) for x in xs]
        torch.cuda.synchronize()
        gms_b, gms_1 = timed(g_b.replay), timed(g_1.replay)
        same = all(torch.equal(a, b) for a, b in zip(keep_b, keep_1))
        print(json.dumps({"n": n, "matrices": B, "nbits": nbits, "MB": round(by / 1e6, 1),
                          "batched_ms": round(gms_b, 4), "batched_GBs": round(by / gms_b / 1e6, 1), "batched_frac_hbm": round(by / gms_b / 1e6 / HBM, 3),
                          "per_matrix_ms": round(gms_1, 4), "per_matrix_GBs": round(by / gms_1 / 1e6, 1), "per_matrix_frac_hbm": round(by / gms_1 / 1e6 / HBM, 3),
                          "eager_batched_ms": round(ms_b, 3), "eager_per_matrix_ms": round(ms_1, 3), "same": same}), flush=True)
        del xs, keep_b, keep_1, g_b, g_1

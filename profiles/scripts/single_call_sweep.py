#!/usr/bin/env python3
"""Latency of ONE bitMM2Bit call (1-bit adjacency n x n, b-bit features n x N) per engine, CUDA-graph replay of 20
calls so the host is out of the picture: where should QGTC_ENGINE_AUTO switch from mma.sync to tcgen05?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402


def bench(engine, n, N, b, reps=20):
    a = capi.val2bit((torch.rand((n, n), device="cuda") < 0.02).float(), 1, False, False)
    x = capi.val2bit(torch.randint(0, 1 << b, (n, N), device="cuda").float(), b, True, False)
    for _ in range(3):
        capi.bitMM2Bit(a, x, n, n, N, 1, b, b, engine=engine)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                capi.bitMM2Bit(a, x, n, n, N, 1, b, b, engine=engine)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


if __name__ == "__main__":
    for n in (759, 1283, 2258, 4096, 8192):
        for N in (16, 64):
            r = {"n": n, "N": N, "bits": 2, "gop": round(2.0 * n * n * N / 1e9, 3)}
            for name, e in (("mma", capi.ENGINE_MMA_SYNC), ("tcgen05", capi.ENGINE_TCGEN05), ("grouped", capi.ENGINE_GROUPED), ("auto", capi.ENGINE_AUTO)):
                try:
                    r[name + "_us"] = round(bench(e, n, N, 2), 2)
                except Exception as ex:
                    r[name + "_us"] = str(ex)[:60]
            print(json.dumps(r), flush=True)

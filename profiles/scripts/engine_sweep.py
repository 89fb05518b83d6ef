#!/usr/bin/env python3
"""Engine A/B measurement (DESIGN.md "tensor path chosen by measurement"): batched bit-GEMM TOPS of
the mma.sync and tcgen05 engines over the Fig. 8a shapes, CUDA events, inputs larger than L2 for the
4096 shapes.  Prints one JSON line per (engine, M, N, bits)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402


def run(engine, M, K, N, b, G, reps=5, flags=0):
    wa = capi.packed_words(0, 1, M, K)
    A = torch.randint(-2**31, 2**31 - 1, (G, wa), dtype=torch.int64, device="cuda").to(torch.int32)
    X = torch.empty((G, capi.packed_words(1, b, K, N)), dtype=torch.int32, device="cuda")
    for g in range(G):
        X[g] = capi.val2bit(torch.randint(0, 1 << b, (K, N), device="cuda").float(), b, True, False).view(-1)
    out = torch.empty((G, capi.packed_words(0, b, M, N)), dtype=torch.int32, device="cuda")
    for _ in range(3):
        capi.bitMM2Bit_batched(A, X, G, M, K, N, 1, b, b, out=out, engine=engine, flags=flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        capi.bitMM2Bit_batched(A, X, G, M, K, N, 1, b, b, out=out, engine=engine, flags=flags)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * M * K * N * G / (ms * 1e-3) / 1e12


if __name__ == "__main__":
    engines = {"mma": capi.ENGINE_MMA_SYNC, "tcgen05": capi.ENGINE_TCGEN05}
    shapes = [(4096, 64), (4096, 32), (4096, 16), (2048, 64), (1024, 64), (4096, 128), (4096, 256)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        shapes = [(4096, 64), (4096, 16)]
    for name, e in engines.items():
        for (M, N) in shapes:
            G = max(8, (256 << 20) // (M * M // 8))     # >= 256 MiB of packed adjacency
            G = min(G, 512)
            for b in (1, 2, 4, 8):
                try:
                    ms, tops = run(e, M, M, N, b, G, flags=int(os.environ.get("QGTC_SWEEP_FLAGS", "0"), 0))
                    print(json.dumps({"engine": name, "M": M, "K": M, "N": N, "bits": b, "graphs": G, "ms": round(ms, 4), "tops": round(tops, 1)}), flush=True)
                except Exception as ex:
                    print(json.dumps({"engine": name, "M": M, "N": N, "bits": b, "error": str(ex)[:200]}), flush=True)

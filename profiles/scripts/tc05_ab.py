#!/usr/bin/env python3
"""A/B timing of gemm_tc05_kernel alone (CUDA events around each launch, qgtc_debug_tc05_time) over the Fig. 8a shapes.
Env knobs are read once per process, so run one process per configuration.  One JSON line per (M, N, bits)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402


def run(M, K, N, b, G, reps=6):
    L = capi.lib()
    L.qgtc_debug_tc05_time.argtypes = [ctypes.c_int]
    L.qgtc_debug_tc05_kernel_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    wa = capi.packed_words(0, 1, M, K)
    A = torch.randint(-2**31, 2**31 - 1, (G, wa), dtype=torch.int64, device="cuda").to(torch.int32)
    X = torch.empty((G, capi.packed_words(1, b, K, N)), dtype=torch.int32, device="cuda")
    for g in range(G):
        X[g] = capi.val2bit(torch.randint(0, 1 << b, (K, N), device="cuda").float(), b, True, False).view(-1)
    out = torch.empty((G, capi.packed_words(0, b, M, N)), dtype=torch.int32, device="cuda")
    for _ in range(3):
        capi.bitMM2Bit_batched(A, X, G, M, K, N, 1, b, b, out=out, engine=capi.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    L.qgtc_debug_tc05_time(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        capi.bitMM2Bit_batched(A, X, G, M, K, N, 1, b, b, out=out, engine=capi.ENGINE_TCGEN05)
    e1.record()
    torch.cuda.synchronize()
    L.qgtc_debug_tc05_time(0)
    kt = np.zeros(64, dtype=np.float32)
    nk = L.qgtc_debug_tc05_kernel_times(kt.ctypes.data, 64)
    kms = float(np.median(kt[:nk])) if nk else float("nan")
    ops = 2.0 * M * K * N * G
    return e0.elapsed_time(e1) / reps, kms, ops / (kms * 1e-3) / 1e12


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    shapes = [(4096, 64), (4096, 16), (1024, 64)] if (len(sys.argv) > 2 and sys.argv[2] == "quick") else [(4096, 64), (4096, 32), (4096, 16), (2048, 64), (1024, 64), (4096, 128)]
    for (M, N) in shapes:
        G = min(512, max(8, (192 << 20) // (M * M // 8)))
        for b in (1, 2, 4, 8):
            ms, kms, tops = run(M, M, N, b, G)
            print(json.dumps({"tag": tag, "M": M, "N": N, "bits": b, "graphs": G, "call_ms": round(ms, 4), "kernel_ms": round(kms, 4), "kernel_tops": round(tops, 1)}), flush=True)

#!/usr/bin/env python3
"""On-box library ceiling for the roofline: dense INT8 GEMM (s8 x s8 -> s32) through cuBLASLt (torch._int_mm),
8192^3 and 16384 x 8192 x 8192, best of 10 (burst) and back to back for 2 s (sustained), plus the same for bf16
(the MEASURED_PEAKS.json method).  SURVEY 8(f) rank 4: the cuBLAS comparison point of the reference's Fig. 8a."""
import json
import time

import torch


def bench(fn, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    n = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 2.0:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return best, flops * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


if __name__ == "__main__":
    for (M, K, N) in ((8192, 8192, 8192), (16384, 8192, 8192)):
        a8 = torch.randint(-8, 8, (M, K), dtype=torch.int8, device="cuda")
        b8 = torch.randint(-8, 8, (N, K), dtype=torch.int8, device="cuda").t()      # column-major B, as cuBLASLt wants it
        burst, sus = bench(lambda: torch._int_mm(a8, b8), 2.0 * M * K * N)
        print(json.dumps({"op": "cuBLASLt int8 (torch._int_mm)", "M": M, "K": K, "N": N, "tops_burst": round(burst, 1), "tops_sustained": round(sus, 1)}), flush=True)
        a = torch.randn(M, K, dtype=torch.bfloat16, device="cuda"); b = torch.randn(K, N, dtype=torch.bfloat16, device="cuda")
        burst, sus = bench(lambda: torch.matmul(a, b), 2.0 * M * K * N)
        print(json.dumps({"op": "cuBLAS bf16 (torch.matmul)", "M": M, "K": K, "N": N, "tflops_burst": round(burst, 1), "tflops_sustained": round(sus, 1)}), flush=True)

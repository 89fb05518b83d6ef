#!/usr/bin/env python3
"""cuBLASLt INT8 (torch._int_mm, s8 x s8 -> s32) at the Fig. 8a shapes of the reference's cuBLAS table
(cuBLASGemmEX/cublas_main.cu:123-172: M = K in {1024, 2048, 4096}, N in {16, 32, 64}), 200 back-to-back calls between
two CUDA events as there, beside this repo's bitMM2Bit on the same shapes (one GEMM per launch, 2-bit features, the
reference's own protocol QGTC_device.cu:403-418) and its batched form.  TOPS = 2*M*K*N / t."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

REPS = 200


def timed(fn, reps=REPS):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def graphed(fn, reps=REPS):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


if __name__ == "__main__":
    for M in (1024, 2048, 4096):
        for N in (16, 32, 64):
            K = M
            ops = 2.0 * M * K * N
            a8 = torch.randint(0, 2, (M, K), dtype=torch.int8, device="cuda")
            b8 = torch.randint(0, 4, (N, K), dtype=torch.int8, device="cuda").t()
            r = {"M": M, "K": K, "N": N}
            try:
                r["cublaslt_int8_tops_eager"] = round(ops / timed(lambda: torch._int_mm(a8, b8)) / 1e12, 1)
                r["cublaslt_int8_tops_graph"] = round(ops / graphed(lambda: torch._int_mm(a8, b8)) / 1e12, 1)
            except Exception as e:  # noqa: BLE001
                r["cublaslt_error"] = str(e)[:120]
            A = capi.val2bit(a8.float(), 1, False, False)
            X = capi.val2bit(b8.t().contiguous().float(), 2, True, False)
            r["qgtc_b200_single_call_tops_eager"] = round(ops / timed(lambda: capi.bitMM2Bit(A, X, M, K, N, 1, 2, 2)) / 1e12, 1)
            r["qgtc_b200_single_call_tops_graph"] = round(ops / graphed(lambda: capi.bitMM2Bit(A, X, M, K, N, 1, 2, 2)) / 1e12, 1)
            t = timed(lambda: capi.bitMM2Bit_aliased(A, X, REPS, M, K, N, 1, 2, 2), reps=5) / REPS
            r["qgtc_b200_profile_convention_tops"] = round(ops / t / 1e12, 1)
            print(json.dumps(r), flush=True)

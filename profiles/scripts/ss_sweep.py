#!/usr/bin/env python3
"""u8 operand path of the tcgen05 batch engine with and without the shared-memory fourth stage (QGTC_TC05_SS, read once per
process: run this script once per setting).  Batched 4096x4096x64 (and x16 / x32), 96-128 graphs, CUDA events around the calls."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from qgtc_ppopp22_b200 import capi  # noqa: E402
from engine_sweep import run  # noqa: E402

for (M, N) in ((4096, 64), (4096, 32), (2048, 64), (1024, 64)):
    G = min(512, max(8, (256 << 20) // (M * M // 8)))
    for b in (1, 2, 4, 8):
        ms, tops = run(capi.ENGINE_TCGEN05, M, M, N, b, G)
        print(json.dumps({"ss": os.environ.get("QGTC_TC05_SS", "0"), "fp4": os.environ.get("QGTC_TC05_FP4", "1"), "M": M, "N": N, "bits": b, "graphs": G,
                          "ms": round(ms, 4), "tops": round(tops, 1)}), flush=True)

#!/usr/bin/env python3
"""Where does the tcgen05 engine's time go?  Runs the batched 4096x4096xN GEMM with individual pipeline
stages disabled through the library's timing-experiment flag bits (results are then wrong; timing only)."""
import os
os.environ.setdefault("QGTC_TC05_DEBUG_KNOBS", "1")   # the library refuses the knob bits otherwise
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from engine_sweep import run  # noqa: E402

KNOBS = {"none": 0, "dbg_instance": 0x200000, "alt_accumulators": 0x400000, "nanosleep_waits": 0x200000, "nojump": 1, "no_st": 0x10000, "no_mma": 0x20000, "no_bulk": 0x40000, "no_ldg": 0x80000 | 1,
         "no_st+no_mma": 0x30000, "no_st+no_mma+no_bulk": 0x70000, "all_off": 0xF0000 | 1}

if __name__ == "__main__":
    for N in (64, 16):
        for name, f in KNOBS.items():
            ms, tops = run(capi.ENGINE_TCGEN05, 4096, 4096, N, 2, 128, flags=f)
            print(json.dumps({"N": N, "knob": name, "ms": round(ms, 4), "tops": round(tops, 1)}), flush=True)

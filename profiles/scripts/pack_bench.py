#!/usr/bin/env python3
"""val2bit (quantize-and-pack) bandwidth: algorithmic bytes = 4*H*W read + packed bytes written, per call;
inputs rotate through > L2 worth of distinct matrices.  Prints JSON lines with GB/s and the fraction of the
measured HBM copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0


def bench(H, W, bits, col_major, copies, reps=5):
    xs = [torch.rand((H, W), device="cuda") * (1 << bits) for _ in range(copies)]
    out = None
    for x in xs[:2]:
        out = capi.val2bit(x, bits, col_major, False)
    torch.cuda.synchronize()
    # one CUDA graph over all copies: the python/ctypes call overhead (~15 us) must not starve the GPU
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for x in xs:
                out = capi.val2bit(x, bits, col_major, False)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * copies)
    nbytes = 4.0 * H * W + out.numel() * 4
    gbs = nbytes / (ms * 1e-3) / 1e9
    return ms, gbs


if __name__ == "__main__":
    for (H, W, col, copies) in [(4096, 4096, False, 8), (8192, 8192, False, 3), (2258, 2258, False, 32), (759, 759, False, 64),
                                (4096, 64, True, 64), (4096, 4096, True, 8)]:
        for bits in ((1, 2, 8) if not col else (2, 8)):
            ms, gbs = bench(H, W, bits, col, copies)
            print(json.dumps({"op": "val2bit", "H": H, "W": W, "bits": bits, "col_major": col, "us": round(ms * 1e3, 2),
                              "GBs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / HBM, 3)}), flush=True)

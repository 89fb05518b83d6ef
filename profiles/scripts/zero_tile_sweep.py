#!/usr/bin/env python3
"""BASELINE configs[4] (Fig. 8b / 9): zero-tile jumping and adjacency-size sweep.  Synthetic sparse
adjacencies (i.i.d. density p, and block-diagonal with blocks of 128) at partition sizes 1K-16K, N = 64,
2-bit features: TOPS (dense-equivalent 2*M*K*N) with and without jumping, non-zero tile ratio as the
reference's counters report it (8 x 128-bit tiles, parse_counter.py:16)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def make_adj(M, kind, p, gen):
    if kind == "iid":
        return (torch.rand((M, M), device="cuda", generator=gen) < p).float()
    a = torch.zeros((M, M), device="cuda")
    for b0 in range(0, M, 128):
        blk = (torch.rand((128, 128), device="cuda", generator=gen) < p).float()
        a[b0:b0 + 128, b0:b0 + 128] = blk[: M - b0, : M - b0]
    return a


if __name__ == "__main__":
    N, b = 64, 2
    gen = torch.Generator(device="cuda").manual_seed(3)
    for M in (1024, 2048, 4096, 8192, 16384):
        G = max(2, min(64, (256 << 20) // (M * M // 8)))
        X = torch.stack([capi.val2bit(torch.randint(0, 4, (M, N), device="cuda", generator=gen).float(), b, True, False).view(-1) for _ in range(G)])
        for kind, p in (("iid", 1.0), ("iid", 0.1), ("iid", 1e-2), ("iid", 1e-3), ("iid", 1e-4), ("blockdiag128", 0.5)):
            A = torch.stack([capi.val2bit(make_adj(M, kind, p, gen), 1, False, False).view(-1) for _ in range(G)])
            total, nz = capi.tile_counts(A[0].view(-1, M // 32 if M % 128 == 0 else -1), M, M, N, 1, b)
            out = torch.empty((G, capi.packed_words(0, b, M, N)), dtype=torch.int32, device="cuda")
            res = {"M": M, "N": N, "bits": b, "pattern": kind, "density": p, "graphs": G, "nonzero_tile_ratio": round(nz / total, 4)}
            bm = capi.tile_bitmap(A, M, M, G)
            ms = timed(lambda: capi.bitMM2Bit_tilemap(A, X, bm, G, M, M, N, b, b, out=out, engine=capi.ENGINE_TCGEN05))
            res["tcgen05_tilemap_tops"] = round(2.0 * M * M * N * G / (ms * 1e-3) / 1e12, 1)
            res["tile128_nonzero_ratio"] = round(float(sum(bin(int(v) & 0xffffffff).count("1") for v in bm.cpu().numpy().reshape(-1).tolist())) / (G * ((M + 127) // 128) ** 2), 4)
            for eng_name, eng in (("tcgen05", capi.ENGINE_TCGEN05), ("mma", capi.ENGINE_MMA_SYNC)):
                for jname, fl in (("jump", 0), ("nojump", capi.FLAG_NO_ZERO_JUMP)):
                    ms = timed(lambda: capi.bitMM2Bit_batched(A, X, G, M, M, N, 1, b, b, out=out, engine=eng, flags=fl))
                    res["%s_%s_tops" % (eng_name, jname)] = round(2.0 * M * M * N * G / (ms * 1e-3) / 1e12, 1)
            print(json.dumps(res), flush=True)
            del A

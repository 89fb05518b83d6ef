#!/usr/bin/env python3
"""Reduced parity set for compute-sanitizer (memcheck / racecheck / synccheck): the tcgen05 engine's ragged, batched,
tile-bitmap, pair and 4-bit-operand cases, the mma.sync engine, the pack kernels and (when built) the grouped engine.
Every result is still compared with the mma.sync engine / a dense torch reference, so a sanitizer run is also a
parity run.  Usage:  compute-sanitizer --tool racecheck python profiles/scripts/sanitize_subset.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

torch.manual_seed(7)
bad = []


def ops(M, K, N, b, density=0.2):
    A = (torch.rand(M, K, device="cuda") < density).float()
    X = torch.randint(0, 1 << b, (K, N), device="cuda").float()
    return A, X, capi.val2bit(A, 1, False, False), capi.val2bit(X, b, True, False)


# single problems: ragged tiles, short K, N up to 256, every operand path (fp4: b <= 2, u8 mode 2: b <= 5, mode 1: b = 8)
for (M, K, N, b) in [(130, 513, 16, 1), (300, 1000, 40, 2), (257, 640, 72, 4), (264, 1100, 200, 8), (8, 128, 8, 2)]:
    A, X, a, x = ops(M, K, N, b)
    want = (A.double() @ X.double()).float()
    got = capi.bitMM2Int(a, x, M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05)
    if not torch.equal(got, want):
        bad.append(("int", M, K, N, b))
    ref = capi.bitMM2Bit(a, x, M, K, N, 1, b, b, engine=capi.ENGINE_MMA_SYNC)
    for fl in (0, capi.FLAG_NO_ZERO_JUMP, capi.FLAG_TC05_PAIR):
        if not torch.equal(capi.bitMM2Bit(a, x, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05, flags=fl), ref):
            bad.append(("bits", M, K, N, b, fl))
    bm = capi.tile_bitmap(a, M, K)
    if not torch.equal(capi.bitMM2Bit_tilemap(a.view(1, -1), x.view(1, -1), bm, 1, M, K, N, b, b, engine=capi.ENGINE_TCGEN05).view(-1), ref.view(-1)):
        bad.append(("tilemap", M, K, N, b))

# many short ragged problems in one launch (the case that exposed the issuer phase race in round 1)
for (M, K, N, b, cnt) in [(130, 513, 64, 2, 40), (264, 1100, 16, 1, 40), (129, 385, 128, 4, 30)]:
    A, X, a, x = ops(M, K, N, b, 0.3)
    ref = capi.bitMM2Bit(a, x, M, K, N, 1, b, b, engine=capi.ENGINE_MMA_SYNC)
    out = capi.bitMM2Bit_batched(a, x, cnt, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
    if not bool((out == ref.view(1, -1)).all()):
        bad.append(("batched", M, K, N, b, cnt))

if hasattr(capi, "sanitize_extra"):
    bad += capi.sanitize_extra()
torch.cuda.synchronize()
print("SANITIZE_SUBSET", "BAD %r" % bad if bad else "OK")
sys.exit(1 if bad else 0)

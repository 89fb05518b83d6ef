#!/usr/bin/env python3
"""Randomised parity sweep of the tcgen05 batch engine against the mma.sync engine (both through the C ABI): random shapes, bit
widths, batch sizes, densities, with and without zero-tile jumping and tile bitmaps.  A protocol race shows up as a rare mismatch,
so every configuration is repeated.  Usage: python profiles/scripts/fuzz_tc05.py [seconds] [seed]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import capi  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(seed)
torch.manual_seed(seed)


def ri(lo, hi):
    return int(torch.randint(lo, hi + 1, (1,), generator=g))


t0 = time.time()
n_cfg = n_bad = n_unsup = 0
while time.time() - t0 < budget:
    kind = ri(0, 9)
    if kind < 4:      # headline-like: compile-time instances, several tiles per CTA
        M, K = 128 * ri(1, 40), 128 * ri(7, 40)
        N = [16, 32, 64][ri(0, 2)]
    elif kind < 8:    # ragged
        M, K, N = ri(1, 3000), ri(1, 5000), ri(1, 256)
    else:             # short K / tiny
        M, K, N = ri(1, 600), ri(1, 300), ri(1, 64)
    b = [1, 2, 3, 4, 5, 8][ri(0, 5)]
    cnt = ri(1, 6) if M * K > 2_000_000 else ri(1, 48)
    dens = [0.5, 0.02, 0.001][ri(0, 2)]
    A = (torch.rand(M, K, device="cuda") < dens).float()
    if ri(0, 3) == 0:   # block-diagonal: whole zero tiles
        blk = torch.zeros_like(A)
        s = 128
        for i in range(0, min(M, K), s):
            blk[i:i + s, i:i + s] = A[i:i + s, i:i + s]
        A = blk
    X = torch.randint(0, 1 << b, (K, N), device="cuda").float()
    a, x = capi.val2bit(A, 1, False, False), capi.val2bit(X, b, True, False)
    ob = [b, 1, 8][ri(0, 2)]
    ref = capi.bitMM2Bit(a, x, M, K, N, 1, b, ob, engine=capi.ENGINE_MMA_SYNC)
    ok = True
    try:
        for rep in range(3):
            fl = [0, capi.FLAG_NO_ZERO_JUMP][rep & 1]
            out = capi.bitMM2Bit_batched(a, x, cnt, M, K, N, 1, b, ob, engine=capi.ENGINE_TCGEN05, flags=fl)
            ok = ok and bool((out == ref.view(1, -1)).all())
        bm = capi.tile_bitmap(a, M, K)
        ok = ok and torch.equal(capi.bitMM2Bit_tilemap(a.view(1, -1), x.view(1, -1), bm, 1, M, K, N, b, ob, engine=capi.ENGINE_TCGEN05).view(-1), ref.view(-1))
        want = (A.double() @ X.double()).float()
        ok = ok and torch.equal(capi.bitMM2Int(a, x, M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05), want)
    except capi.QgtcError as e:
        n_unsup += 1
        if "unsupported" not in str(e).lower() or n_unsup <= 3:
            ok = ok and "unsupported" in str(e).lower()
            print("error", M, K, N, b, ob, cnt, e, flush=True)
    n_cfg += 1
    if not ok:
        n_bad += 1
        print("MISMATCH", dict(M=M, K=K, N=N, bits=b, out_bit=ob, count=cnt, density=dens), flush=True)
print("FUZZ_TC05 configs %d (refused as unsupported: %d) mismatches %d seconds %.0f" % (n_cfg, n_unsup, n_bad, time.time() - t0))
sys.exit(1 if n_bad else 0)

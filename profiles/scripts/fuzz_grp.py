#!/usr/bin/env python3
"""Randomised parity sweep of the grouped engine (qgtc_bitmm_grouped / qgtc_bitmm_grouped_fused) against per-problem calls of the
mma.sync engine: random problem lists (variable M and K), bit widths on both sides, the three epilogues, tile bitmaps, and the fused
aggregation + update launch against its two separate ops.  Usage: python profiles/scripts/fuzz_grp.py [seconds] [seed]"""
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


fails = {}


def chk(ok, cond, what, m, k):
    if not cond:
        key = what.split()[0]
        fails[key] = fails.get(key, 0) + 1
        if fails[key] <= 4:
            print("  fail:", what, "M", m, "K", k, flush=True)
    return ok and cond


t0 = time.time()
n_cfg = n_bad = n_fused = n_unsup = 0
while time.time() - t0 < budget:
    cnt = ri(1, 40)
    adj = ri(0, 2) > 0                  # adjacency product (1-bit A, K = M) or weight product (b-bit A, small K)
    b1 = 1 if adj else [1, 2, 4, 8][ri(0, 3)]
    b2 = [1, 2, 3, 4, 8][ri(0, 4)]
    N = [8, 16, 40, 64, 121, 128][ri(0, 5)]
    Ms = [ri(1, 2400) for _ in range(cnt)]
    Ks = list(Ms) if adj else [ri(1, 200)] * cnt
    x1s, x2s, As, Xs = [], [], [], []
    for m, k in zip(Ms, Ks):
        A = (torch.rand(m, k, device="cuda") < [0.3, 0.01][ri(0, 1)]).float() if b1 == 1 else torch.randint(0, 1 << b1, (m, k), device="cuda").float()
        X = torch.randint(0, 1 << b2, (k, N), device="cuda").float()
        As.append(A); Xs.append(X)
        x1s.append(capi.val2bit(A, b1, False, False)); x2s.append(capi.val2bit(X, b2, True, False))
    ob = [b2, 1, 8][ri(0, 2)]
    maps = [capi.tile_bitmap(a, m, k).view(-1) for a, m, k in zip(x1s, Ms, Ks)] if (b1 == 1 and ri(0, 1)) else None
    ok = True
    try:
        got = capi.bitmm_grouped(x1s, x2s, Ms, Ks, N, b1, b2, ob, kind=capi.GROUP_BITS_ROW, tile_maps=maps, engine=capi.ENGINE_GROUPED)
        for a, x, m, k, o in zip(x1s, x2s, Ms, Ks, got):
            ok = chk(ok, torch.equal(o, capi.bitMM2Bit(a, x, m, k, N, b1, b2, ob, engine=capi.ENGINE_MMA_SYNC)), 'row', m, k)
        gotf = capi.bitmm_grouped(x1s, x2s, Ms, Ks, N, b1, b2, ob, kind=capi.GROUP_FLOAT, pad_128=True, tile_maps=maps, engine=capi.ENGINE_GROUPED)   # x2 is packed C128: the PAD128 form of bitMM2Int
        for A, X, o in zip(As, Xs, gotf):
            ok = chk(ok, torch.equal(o, (A.double() @ X.double()).float()), 'float', A.shape[0], A.shape[1])
        gotc = capi.bitmm_grouped(x1s, x2s, Ms, Ks, N, b1, b2, ob, kind=capi.GROUP_BITS_COL, engine=capi.ENGINE_GROUPED)
        for a, x, m, k, o in zip(x1s, x2s, Ms, Ks, gotc):
            ok = chk(ok, torch.equal(o, capi.bitMM2Bit_col(a, x, m, k, N, b1, b2, ob, engine=capi.ENGINE_MMA_SYNC)), 'col', m, k)
        if adj and N <= 64:       # fused aggregation + update against the two ops
            N2 = [16, 64, 121][ri(0, 2)]
            bw = [1, 2, 4][ri(0, 2)]
            W = capi.val2bit(torch.randint(0, 1 << bw, (N, N2), device="cuda").float(), bw, True, False)
            ob2 = [ob, 2][ri(0, 1)]
            try:
                f2, f1 = capi.bitmm_grouped_fused(x1s, x2s, Ms, Ks, N, 1, b2, ob, W, N2, bw, ob2, tile_maps=maps, write_out1=True, engine=capi.ENGINE_GROUPED)
                n_fused += 1
                for a, x, m, k, o1, o2 in zip(x1s, x2s, Ms, Ks, f1, f2):
                    r1 = capi.bitMM2Bit(a, x, m, k, N, 1, b2, ob, engine=capi.ENGINE_MMA_SYNC)
                    ok = chk(ok, torch.equal(o1, r1), 'fused-o1', m, k)
                    ok = chk(ok, torch.equal(o2, capi.bitMM2Bit(r1, W, m, N, N2, ob, bw, ob2, engine=capi.ENGINE_MMA_SYNC)), 'fused-o2 N2=%d bw=%d ob2=%d' % (N2, bw, ob2), m, k)
            except capi.QgtcError as e:
                if "unsupported" not in str(e).lower():
                    raise
    except capi.QgtcError as e:
        if "unsupported" in str(e).lower():     # shapes the grouped kernel refuses when it is forced (AUTO falls back to per-problem calls)
            n_unsup += 1
        else:
            ok = False
            print("error", cnt, b1, b2, N, ob, e, flush=True)
    n_cfg += 1
    if not ok:
        n_bad += 1
        print("MISMATCH", dict(count=cnt, bit1=b1, bit2=b2, N=N, out_bit=ob, adj=adj, maps=maps is not None, Ms=Ms[:6]), flush=True)
print("failures by check:", fails)
print("FUZZ_GRP lists %d (fused launches %d, lists refused as unsupported %d) mismatches %d seconds %.0f" % (n_cfg, n_fused, n_unsup, n_bad, time.time() - t0))
sys.exit(1 if n_bad else 0)

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
# (257 x 640 x 64 at 4 bits and 264 x 1100 x 48 at 8 bits run the compile-time four-chunk instances of the u8 path, 300 x 1000 x 40 at 2 bits the seven-chunk 4-bit instance)
for (M, K, N, b) in [(130, 513, 16, 1), (300, 1000, 40, 2), (257, 640, 72, 4), (264, 1100, 200, 8), (8, 128, 8, 2), (257, 640, 64, 4), (264, 1100, 48, 8)]:
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
for (M, K, N, b, cnt) in [(130, 513, 64, 2, 40), (264, 1100, 16, 1, 40), (129, 385, 128, 4, 30), (200, 1024, 64, 4, 40), (136, 1920, 32, 8, 24)]:
    A, X, a, x = ops(M, K, N, b, 0.3)
    ref = capi.bitMM2Bit(a, x, M, K, N, 1, b, b, engine=capi.ENGINE_MMA_SYNC)
    out = capi.bitMM2Bit_batched(a, x, cnt, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
    if not bool((out == ref.view(1, -1)).all()):
        bad.append(("batched", M, K, N, b, cnt))

# grouped engine: multi-plane A, all epilogues, variable shapes in one launch, tile bitmaps, shared-weight stage re-use
for (M, K, N, b1, b2) in [(130, 513, 16, 1, 2), (300, 200, 40, 2, 2), (257, 129, 72, 4, 4), (64, 16, 16, 8, 8)]:
    A = (torch.rand(M, K, device="cuda") < 0.2).float() if b1 == 1 else torch.randint(0, 1 << b1, (M, K), device="cuda").float()
    X = torch.randint(0, 1 << b2, (K, N), device="cuda").float()
    a, x = capi.val2bit(A, b1, False, False), capi.val2bit(X, b2, True, False)
    if not torch.equal(capi.bitMM2Int(a, x, M, K, N, b1, b2, True, engine=capi.ENGINE_GROUPED), (A.double() @ X.double()).float()):
        bad.append(("grp int", M, K, N, b1, b2))
    if not torch.equal(capi.bitMM2Bit(a, x, M, K, N, b1, b2, b2, engine=capi.ENGINE_GROUPED), capi.bitMM2Bit(a, x, M, K, N, b1, b2, b2, engine=capi.ENGINE_MMA_SYNC)):
        bad.append(("grp bits", M, K, N, b1, b2))
    if not torch.equal(capi.bitMM2Bit_col(a, x, M, K, N, b1, b2, b2, engine=capi.ENGINE_GROUPED), capi.bitMM2Bit_col(a, x, M, K, N, b1, b2, b2, engine=capi.ENGINE_MMA_SYNC)):
        bad.append(("grp col", M, K, N, b1, b2))
shapes = [(759, 759), (128, 128), (1, 5), (300, 2000), (130, 129)] * 60
pool = {}
for (M, K) in set(shapes):
    A = (torch.rand(M, K, device="cuda") < 0.05).float()
    X = torch.randint(0, 4, (K, 16), device="cuda").float()
    a, x = capi.val2bit(A, 1, False, False), capi.val2bit(X, 2, True, False)
    pool[(M, K)] = (a, x, capi.tile_bitmap(a, M, K).view(-1), capi.bitMM2Bit(a, x, M, K, 16, 1, 2, 2, engine=capi.ENGINE_MMA_SYNC))
for maps in (None, [pool[s][2] for s in shapes]):
    outs = capi.bitmm_grouped([pool[s][0] for s in shapes], [pool[s][1] for s in shapes], [s[0] for s in shapes], [s[1] for s in shapes],
                              16, 1, 2, 2, tile_maps=maps, engine=capi.ENGINE_GROUPED)
    if not all(torch.equal(o, pool[s][3]) for o, s in zip(outs, shapes)):
        bad.append(("grp list", maps is not None))
W = capi.val2bit(torch.ones(50, 16, device="cuda"), 2, True, False)
xs = [capi.val2bit(torch.randint(0, 4, (200 + i, 50), device="cuda").float(), 2, False, False) for i in range(8)] * 80
outs = capi.bitmm_grouped(xs, [W] * len(xs), [200 + (i % 8) for i in range(len(xs))], [50] * len(xs), 16, 2, 2, 2, engine=capi.ENGINE_GROUPED)
for i in range(8):
    if not torch.equal(outs[i + 8 * 79], capi.bitMM2Bit(xs[i], W, 200 + i, 50, 16, 2, 2, 2, engine=capi.ENGINE_MMA_SYNC)):
        bad.append(("grp reuse", i))
# pack kernels, all layouts
for (H, W_, nb) in [(37, 300, 1), (130, 513, 3), (257, 4100, 8), (759, 50, 2)]:
    xf = torch.rand(H, W_, device="cuda") * (1 << nb) * 1.2 - 0.3
    for cm, ol in [(False, False), (True, False), (True, True)]:
        back = capi.bit2val(capi.val2bit(xf, nb, cm, ol), nb, H, W_, cm, ol)
        if not (cm and ol):
            y = torch.where(xf < 0, torch.ones_like(xf), torch.where(xf > (1 << nb), torch.full_like(xf, (1 << nb) - 1), xf))
            if not torch.equal(back, torch.round(y).to(torch.int32) & ((1 << nb) - 1)):
                bad.append(("pack", H, W_, nb, cm))
torch.cuda.synchronize()
print("SANITIZE_SUBSET", "BAD %r" % bad if bad else "OK")
sys.exit(1 if bad else 0)

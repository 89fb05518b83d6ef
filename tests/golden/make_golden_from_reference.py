#!/usr/bin/env python3
"""Generate golden vectors by running the REFERENCE's own kernels on a B200.

Run on the GPU box (the extension is prebuilt by oracle/build_ref.py and travels in oracle/_ref/):

    python tests/golden/make_golden_from_reference.py gpurun_out/golden

For every case it stores the seeded inputs and the outputs of the reference extension
(variant QGTC_ref_lw = reference sources + logical warp index, see oracle/ref_logical_warpid.h;
the unmodified QGTC_ref is run as well and its agreement recorded) in one .npz, plus a report.json
saying, per case and op, whether reference == CPU oracle and reference == this repo's CUDA path.
The .npz files are then committed under tests/golden/ and replayed on CPU by tests/test_golden.py.

Only the reference's well-defined domain is exercised: well-formed operands,
C8 packing only where P8(W)==P128(W) or bits==1, no zero-jump kernel (its probe reads out of bounds,
kernel.h:581).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import bitplane_oracle as O  # noqa: E402
import build_ref  # noqa: E402

# (name, M, K, N, bit1, bit2, out_bit, adjacency density or None for multi-bit A)
CASES = [
    ("kat_ones_3", 3, 3, 3, 1, 2, 2, None),
    ("ragged_37x300x21", 37, 300, 21, 1, 3, 3, 0.3),
    ("tile_64x128x16", 64, 128, 16, 1, 1, 1, 0.5),
    ("multi_130x513x70", 130, 513, 70, 2, 2, 2, None),
    ("ppi_batch_759x759x16", 759, 759, 16, 1, 2, 2, 0.02),
    ("xw_759x50x16", 759, 50, 16, 2, 2, 2, None),
    ("classes_200x1000x121", 200, 1000, 121, 2, 2, 2, None),
    ("wide_129x257x33_8bit", 129, 257, 33, 8, 8, 8, None),
    ("sat_255x129x65", 255, 129, 65, 3, 5, 4, None),
]


def features(rng, shape, nbits):
    hi = 1 << nbits
    x = rng.integers(0, hi, size=shape).astype(np.float32)
    m = rng.random(shape)
    x = np.where(m < 0.05, x + 0.5, x)
    x = np.where((m >= 0.05) & (m < 0.08), -rng.random(shape) * 3, x)
    x = np.where((m >= 0.08) & (m < 0.11), hi + rng.random(shape) * 5, x)
    x = np.where((m >= 0.11) & (m < 0.12), np.float32(hi), x)
    return x.astype(np.float32)


def u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


def run_case(ref, name, M, K, N, b1, b2, ob, density):
    rng = np.random.default_rng(sum(map(ord, name)))
    if name.startswith("kat_ones"):
        A, X = np.ones((M, K), np.float32), np.ones((K, N), np.float32)
    else:
        A = (rng.random((M, K)) < density).astype(np.float32) if density is not None else features(rng, (M, K), b1)
        X = features(rng, (K, N), b2)
    Ad, Xd = torch.from_numpy(A).cuda(), torch.from_numpy(X).cuda()
    out = {"A": A, "X": X, "meta": np.array([M, K, N, b1, b2, ob], np.int64)}
    out["a_row"] = u32(ref.val2bit(Ad, b1, False, False))
    out["x_col"] = u32(ref.val2bit(Xd, b2, True, False))
    c8_ok = b2 == 1 or O.P8(N) == O.P128(N)
    if c8_ok:
        out["x_col8"] = u32(ref.val2bit(Xd, b2, True, True))
    ta, tx = torch.from_numpy(out["a_row"].view(np.int32)).cuda(), torch.from_numpy(out["x_col"].view(np.int32)).cuda()
    out["x_unpacked"] = ref.bit2val(tx, b2, K, N, True, False).cpu().numpy()
    out["mm2bit"] = u32(ref.bitMM2Bit(ta, tx, M, K, N, b1, b2, ob))
    out["mm2bit_base_cnt"] = u32(ref.bitMM2Bit_base_cnt(ta, tx, M, K, N, b1, b2, ob))
    out["mm2int_pad128"] = ref.bitMM2Int(ta, tx, M, K, N, b1, b2, True).cpu().numpy()
    if c8_ok:
        tx8 = torch.from_numpy(out["x_col8"].view(np.int32)).cuda()
        out["mm2int_pad8"] = ref.bitMM2Int(ta, tx8, M, K, N, b1, b2, False).cpu().numpy()
    out["mm2bit_col_shipped"] = u32(ref.bitMM2Bit_col(ta, tx, M, K, N, b1, b2, ob))
    torch.cuda.synchronize()
    return out


def oracle_outputs(g):
    M, K, N, b1, b2, ob = [int(v) for v in g["meta"]]
    A, X = g["A"], g["X"]
    o = {"a_row": O.val2bit(A, b1, False, False), "x_col": O.val2bit(X, b2, True, False)}
    if "x_col8" in g:
        o["x_col8"] = O.val2bit(X, b2, True, True)
        o["mm2int_pad8"] = O.bitMM2Int(o["a_row"], o["x_col8"], M, K, N, b1, b2, False)
    o["x_unpacked"] = O.bit2val(o["x_col"], b2, K, N, True)
    o["mm2bit"] = O.bitMM2Bit(o["a_row"], o["x_col"], M, K, N, b1, b2, ob)
    o["mm2bit_base_cnt"] = o["mm2bit"]
    o["mm2int_pad128"] = O.bitMM2Int(o["a_row"], o["x_col"], M, K, N, b1, b2, True)
    o["mm2bit_col_shipped"] = O.bitMM2Bit_col(o["a_row"], o["x_col"], M, K, N, b1, b2, ob, ref_quirk=True)
    return o


def ours_outputs(g):
    from qgtc_ppopp22_b200 import capi
    M, K, N, b1, b2, ob = [int(v) for v in g["meta"]]
    Ad, Xd = torch.from_numpy(g["A"]).cuda(), torch.from_numpy(g["X"]).cuda()
    a, x = capi.val2bit(Ad, b1, False, False), capi.val2bit(Xd, b2, True, False)
    o = {"a_row": u32(a), "x_col": u32(x)}
    if "x_col8" in g:
        x8 = capi.val2bit(Xd, b2, True, True)
        o["x_col8"] = u32(x8)
        o["mm2int_pad8"] = capi.bitMM2Int(a, x8, M, K, N, b1, b2, False).cpu().numpy()
    o["x_unpacked"] = capi.bit2val(x, b2, K, N, True, False).cpu().numpy()
    o["mm2bit"] = u32(capi.bitMM2Bit(a, x, M, K, N, b1, b2, ob))
    o["mm2bit_base_cnt"] = o["mm2bit"]
    o["mm2int_pad128"] = capi.bitMM2Int(a, x, M, K, N, b1, b2, True).cpu().numpy()
    o["mm2bit_col_shipped"] = u32(capi.bitMM2Bit_col(a, x, M, K, N, b1, b2, ob, flags=capi.FLAG_COL_REF_QUIRK))
    return o


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(a.reshape(-1).shape == b.reshape(-1).shape and np.array_equal(a.reshape(-1), b.reshape(-1)))


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    report = {"device": torch.cuda.get_device_name(0), "cases": {}}
    ref_lw = build_ref.load_variant("QGTC_ref_lw")
    try:
        ref_plain = build_ref.load_variant("QGTC_ref")
    except Exception as e:  # pragma: no cover
        ref_plain = None
        report["unmodified_load_error"] = repr(e)
    for case in CASES:
        name = case[0]
        g = run_case(ref_lw, *case)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **g)
        orc, ours = oracle_outputs(g), ours_outputs(g)
        rep = {}
        for k in orc:
            rep[k] = {"ref_eq_oracle": same(g[k], orc[k]), "ref_eq_ours": same(g[k], ours[k])}
        if ref_plain is not None:
            gp = run_case(ref_plain, *case)
            for k in orc:
                rep[k]["unmodified_eq_lw"] = same(gp[k], g[k])
        report["cases"][name] = rep
        print(name, json.dumps(rep))
    with open(os.path.join(outdir, "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    bad = [(c, k) for c, r in report["cases"].items() for k, v in r.items() if not (v["ref_eq_oracle"] and v["ref_eq_ours"])]
    print("MISMATCHES:", bad)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")))

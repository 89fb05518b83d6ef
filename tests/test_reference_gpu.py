"""GPU tests against the reference's own kernels (oracle/_ref, rebuilt for sm_100a from the
unmodified sources; see oracle/build_ref.py).  Skipped when the prebuilt extension did not travel."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def ref():
    import build_ref
    if not os.path.exists(build_ref.so_path("QGTC_ref_lw")):
        pytest.skip("oracle/_ref/QGTC_ref_lw not built")
    return build_ref.load_variant("QGTC_ref_lw")


def test_reference_kernels_agree_with_oracle_and_cuda_path(ref, capfd):
    import make_golden_from_reference as mg
    bad = []
    for case in mg.CASES:
        g = mg.run_case(ref, *case)
        orc, ours = mg.oracle_outputs(g), mg.ours_outputs(g)
        for k in orc:
            if not mg.same(g[k], orc[k]):
                bad.append((case[0], k, "oracle"))
            if not mg.same(g[k], ours[k]):
                bad.append((case[0], k, "cuda"))
    capfd.readouterr()  # the reference prints from bit2val / bitMM2Bit_col
    assert not bad, bad


def test_committed_golden_vectors_match_this_box(ref, capfd):
    """The vectors in tests/golden/ were produced by the same reference build: regenerate and compare."""
    import glob
    import make_golden_from_reference as mg
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    if not files:
        pytest.skip("no golden vectors committed")
    cases = {c[0]: c for c in mg.CASES}
    for f in files:
        name = os.path.basename(f)[:-4]
        if name not in cases:
            continue
        g0 = np.load(f)
        g1 = mg.run_case(ref, *cases[name])
        for k in g0.files:
            assert mg.same(g0[k], g1[k]), (name, k)
    capfd.readouterr()

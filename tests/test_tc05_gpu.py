"""GPU parity tests of the tcgen05 engine (engine=ENGINE_TCGEN05 forced): tcgen05.mma kind::i8 with the
adjacency expanded into TMEM, against the CPU oracle, bit-exact."""
import numpy as np
import pytest

from conftest import rand_adj, rand_features

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def capi():
    from qgtc_ppopp22_b200 import capi as c
    c.lib()
    return c


def dev_words(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host_u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


CASES = [
    (128, 128, 16, 1, 1), (128, 256, 64, 2, 2), (8, 128, 8, 1, 1), (3, 3, 3, 2, 2), (300, 1000, 40, 2, 2),
    (759, 759, 16, 2, 2), (515, 515, 29, 2, 2), (2258, 2258, 128, 2, 2), (1000, 3000, 121, 4, 4),
    (257, 129, 200, 8, 8), (130, 513, 256, 3, 5), (1024, 1024, 32, 4, 4), (2048, 2048, 64, 8, 8), (640, 4096, 48, 5, 3),
    (257, 640, 200, 2, 2), (384, 1024, 256, 1, 1), (300, 2000, 144, 2, 2), (520, 1500, 96, 1, 2),
]


@pytest.mark.parametrize("M,K,N,b2,ob", CASES)
def test_tc05_bitmm2bit_and_int(capi, coracle, M, K, N, b2, ob):
    rng = np.random.default_rng(M + K + N)
    A = rand_adj(rng, M, K, 0.2, zero_blocks=True)
    X = rand_features(rng, (K, N), b2)
    ba, bx, bx8 = coracle.val2bit(A, 1, False, False), coracle.val2bit(X, b2, True, False), coracle.val2bit(X, b2, True, True)
    da, dx = dev_words(ba), dev_words(bx)
    E = capi.ENGINE_TCGEN05
    got = capi.bitMM2Bit(da, dx, M, K, N, 1, b2, ob, engine=E)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ba, bx, M, K, N, 1, b2, ob))
    nj = capi.bitMM2Bit(da, dx, M, K, N, 1, b2, ob, engine=E, flags=capi.FLAG_NO_ZERO_JUMP)
    assert torch.equal(nj, got)
    gi = capi.bitMM2Int(da, dx, M, K, N, 1, b2, True, engine=E)
    assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(ba, bx, M, K, N, 1, b2, True))
    g8 = capi.bitMM2Int(da, dev_words(bx8), M, K, N, 1, b2, False, engine=E)
    assert np.array_equal(g8.cpu().numpy(), coracle.bitMM2Int(ba, bx8, M, K, N, 1, b2, False))
    # the two engines agree with each other as well
    assert torch.equal(capi.bitMM2Bit(da, dx, M, K, N, 1, b2, ob, engine=capi.ENGINE_MMA_SYNC), got)
    # 2-CTA cluster form (tcgen05.mma.cta_group::2): same bits
    assert torch.equal(capi.bitMM2Bit(da, dx, M, K, N, 1, b2, ob, engine=E, flags=capi.FLAG_TC05_PAIR), got)
    assert torch.equal(capi.bitMM2Int(da, dx, M, K, N, 1, b2, True, engine=E, flags=capi.FLAG_TC05_PAIR), gi)


@pytest.mark.parametrize("b", [1, 2, 4, 8])
def test_tc05_headline_shape(capi, coracle, b):
    M = K = 4096
    N = 64
    rng = np.random.default_rng(40 + b)
    A = rand_adj(rng, M, K, 0.5 if b != 4 else 0.003, zero_blocks=(b == 2))
    X = rand_features(rng, (K, N), b)
    ba, bx = coracle.val2bit(A, 1, False, False), coracle.val2bit(X, b, True, False)
    got = capi.bitMM2Bit(dev_words(ba), dev_words(bx), M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b))


def test_tc05_raw_memory_operands(capi, coracle):
    """A row-major product fed into the column-major slot (main_qgtc.py:148): zero-extension applies."""
    n, h, b = 1500, 16, 2
    rng = np.random.default_rng(8)
    A = rand_adj(rng, n, n, 0.01)
    prev = coracle.val2bit(rand_features(rng, (n, h), b), b, False, False)      # R(b, n, h) used as the B operand
    ba = coracle.val2bit(A, 1, False, False)
    got = capi.bitMM2Bit(dev_words(ba), dev_words(prev), n, n, h, 1, b, b, engine=capi.ENGINE_TCGEN05)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ba, prev, n, n, h, 1, b, b))


def test_tc05_batched(capi, coracle):
    M, K, N, b, cnt = 512, 640, 64, 2, 7
    rng = np.random.default_rng(12)
    x1 = np.stack([coracle.val2bit(rand_adj(rng, M, K, 0.05, zero_blocks=(i % 2 == 0)), 1, False, False).reshape(-1) for i in range(cnt)])
    x2 = np.stack([coracle.val2bit(rand_features(rng, (K, N), b), b, True, False).reshape(-1) for _ in range(cnt)])
    out = capi.bitMM2Bit_batched(dev_words(x1), dev_words(x2), cnt, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
    for i in range(cnt):
        assert np.array_equal(host_u32(out[i]).reshape(-1), coracle.bitMM2Bit(x1[i], x2[i], M, K, N, 1, b, b).reshape(-1)), i
    pair = capi.bitMM2Bit_batched(dev_words(x1), dev_words(x2), cnt, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05, flags=capi.FLAG_TC05_PAIR)
    assert torch.equal(pair, out)


@pytest.mark.parametrize("M,K,N,b,cnt", [(130, 513, 64, 2, 200), (130, 513, 70, 2, 200), (264, 1100, 16, 1, 300), (129, 385, 128, 4, 150)])
def test_tc05_many_short_ragged_problems(capi, coracle, M, K, N, b, cnt):
    """Many small problems per launch with ragged tiles (a few rows in the last tile, one chunk in the last
    stage): the pipeline items alternate between long and very short, which is what exposed an mbarrier
    phase-skipping race between the MMA issuers (a fault under load, never in a single problem)."""
    rng = np.random.default_rng(M * 7 + N)
    a = coracle.val2bit(rand_adj(rng, M, K, 0.3), 1, False, False)
    x = coracle.val2bit(rand_features(rng, (K, N), b), b, True, False)
    want = coracle.bitMM2Bit(a, x, M, K, N, 1, b, b).reshape(-1)
    da, dx = dev_words(a), dev_words(x)
    for _ in range(3):
        out = capi.bitMM2Bit_batched(da, dx, cnt, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
        torch.cuda.synchronize()
        got = host_u32(out)
        assert np.array_equal(got[0].reshape(-1), want) and np.array_equal(got[cnt // 2].reshape(-1), want) and np.array_equal(got[-1].reshape(-1), want)
        assert bool((out == out[0]).all())


def test_tc05_unsupported_shapes_are_refused(capi):
    a = torch.zeros((2 * 8, 4), dtype=torch.int32, device="cuda")
    b = torch.zeros((2 * 4, 128), dtype=torch.int32, device="cuda")
    with pytest.raises(capi.QgtcError, match="unsupported"):
        capi.bitMM2Bit(a, b, 8, 8, 8, 2, 2, 2, engine=capi.ENGINE_TCGEN05)   # multi-bit A goes to the mma engine


@pytest.mark.parametrize("M,K,N,b", [(1024, 1024, 64, 2), (1500, 2100, 16, 1), (4096, 4096, 64, 8), (700, 5000, 200, 4), (130, 16500, 32, 2)])
def test_tile_bitmap_and_tilemap_gemm(capi, oracle, coracle, M, K, N, b):
    """Precomputed tile bitmap (zero-tile jumping): bitmap == oracle, and the GEMM that skips the zero tiles is
    bit-identical to the oracle GEMM."""
    rng = np.random.default_rng(M + N)
    A = np.zeros((M, K), np.float32)
    for b0 in range(0, min(M, K), 128):                      # block-diagonal plus a few stray entries
        A[b0:b0 + 128, b0:b0 + 128] = rng.random((min(128, M - b0), min(128, K - b0))) < 0.3
    A[rng.integers(0, M, 20), rng.integers(0, K, 20)] = 1
    if M > 300:
        A[256:384, :] = 0                                    # one all-zero row tile
    X = rand_features(rng, (K, N), b)
    ba, bx = coracle.val2bit(A, 1, False, False), coracle.val2bit(X, b, True, False)
    da, dx = dev_words(ba), dev_words(bx)
    bm = capi.tile_bitmap(da, M, K)
    want_bm = oracle.tile_bitmap(ba, M, K)
    assert np.array_equal(bm.cpu().numpy().view(np.uint32).reshape(want_bm.shape), want_bm)
    for eng in (capi.ENGINE_TCGEN05, capi.ENGINE_AUTO):
        got = capi.bitMM2Bit_tilemap(da.view(1, -1), dx.view(1, -1), bm, 1, M, K, N, b, b, engine=eng)
        assert np.array_equal(host_u32(got).reshape(-1), coracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b).reshape(-1))
        gi = capi.bitMM2Int_tilemap(da, dx, bm, M, K, N, b, True, engine=eng)
        assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(ba, bx, M, K, N, 1, b, True))


def test_tilemap_batched(capi, coracle):
    M, K, N, b, cnt = 640, 896, 64, 2, 6
    rng = np.random.default_rng(77)
    adjs = []
    for i in range(cnt):
        a = rand_adj(rng, M, K, 0.002 if i % 2 else 0.3, zero_blocks=True)
        if i == 3:
            a[:] = 0                                          # an entirely empty graph
        adjs.append(a)
    x1 = np.stack([coracle.val2bit(a, 1, False, False).reshape(-1) for a in adjs])
    x2 = np.stack([coracle.val2bit(rand_features(rng, (K, N), b), b, True, False).reshape(-1) for _ in range(cnt)])
    d1, d2 = dev_words(x1), dev_words(x2)
    bm = capi.tile_bitmap(d1, M, K, cnt)
    out = capi.bitMM2Bit_tilemap(d1, d2, bm, cnt, M, K, N, b, b, engine=capi.ENGINE_TCGEN05)
    assert torch.equal(capi.bitMM2Bit_tilemap(d1, d2, bm, cnt, M, K, N, b, b, engine=capi.ENGINE_TCGEN05, flags=capi.FLAG_TC05_PAIR), out)
    for i in range(cnt):
        assert np.array_equal(host_u32(out[i]).reshape(-1), coracle.bitMM2Bit(x1[i], x2[i], M, K, N, 1, b, b).reshape(-1)), i


_MODE_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from qgtc_ppopp22_b200 import capi
torch.manual_seed(5)
bad = []
for (M, K, N, b, ob) in [(300, 1000, 40, 2, 2), (515, 515, 64, 4, 4), (257, 640, 72, 5, 3), (1024, 1024, 64, 1, 1), (130, 513, 16, 3, 8)]:
    A = (torch.rand(M, K, device="cuda") < 0.2).float()
    X = torch.randint(0, 1 << b, (K, N), device="cuda").float()
    a, x, x8 = capi.val2bit(A, 1, False, False), capi.val2bit(X, b, True, False), capi.val2bit(X, b, True, True)
    for cnt in (1, 5):
        ref = capi.bitMM2Bit_batched(a, x, cnt, M, K, N, 1, b, ob, engine=capi.ENGINE_MMA_SYNC)
        out = capi.bitMM2Bit_batched(a, x, cnt, M, K, N, 1, b, ob, engine=capi.ENGINE_TCGEN05)
        if not torch.equal(out, ref):
            bad.append(("bits", M, K, N, b, ob, cnt))
    f0 = capi.bitMM2Int(a, x8, M, K, N, 1, b, False, engine=capi.ENGINE_MMA_SYNC)
    f1 = capi.bitMM2Int(a, x8, M, K, N, 1, b, False, engine=capi.ENGINE_TCGEN05)
    if not torch.equal(f0, f1):
        bad.append(("int", M, K, N, b))
print("BAD", bad)
sys.exit(1 if bad else 0)
"""


@pytest.mark.parametrize("mode", [0, 1, 2, "fp4"])
def test_tc05_expansion_modes_agree(mode):
    """The operand paths of the tcgen05 engine -- u8 with the three adjacency expansion modes (plain shift + AND, x128
    multiply + AND, AND-only with pre-scaled feature rows) and the 4-bit kind::mxf4 path for features of at most 2 bits --
    are chosen by operand range; QGTC_TC05_AMODE / QGTC_TC05_FP4 (read once per process) force one where it is exact, so
    each runs in its own interpreter against the mma.sync engine."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QGTC_TC05_FP4="1") if mode == "fp4" else dict(os.environ, QGTC_TC05_AMODE=str(mode), QGTC_TC05_FP4="0")
    r = subprocess.run([sys.executable, "-c", _MODE_SCRIPT.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


# ---------------------------------------------------------------------------------------------------------------
# kind::mxf4 operand path (1- and 2-bit features: E2M1 x E2M1 -> f32 accumulators in TMEM).  The reference accumulates
# in exact int32 (kernel.h:338-342); the f32 accumulation of the tensor core must therefore be exact for every sum the
# launcher's guard (max_sum < 2^23, gemm_tc05.cu launch_gemm_tc05) admits.  These cases drive the accumulators to
# their largest values: all-ones adjacency rows, maximal features, K up to the largest the guard admits.
# ---------------------------------------------------------------------------------------------------------------
def _max_k_under_f4_guard(b):
    gdk = (8388608 - 1) // (128 * ((1 << b) - 1))
    while gdk * 128 * ((1 << b) - 1) >= 8388608:
        gdk -= 1
    return gdk * 128


@pytest.mark.parametrize("b", [1, 2])
@pytest.mark.parametrize("K", [4096, 8192, 16384, 65536, "max"])
@pytest.mark.parametrize("N", [16, 64, 256])
def test_tc05_f4_exact_at_worst_case_sums(capi, b, K, N):
    if K == "max":
        if N != 16:
            pytest.skip("the largest K is run at N = 16 only (operand size)")
        K = _max_k_under_f4_guard(b)
    M = 136                                   # two row tiles, the second one ragged
    top = (1 << b) - 1
    A = torch.ones((M, K), device="cuda")
    da = capi.val2bit(A, 1, False, False)
    del A
    # (1) every feature maximal: every accumulator == K * (2^b - 1)
    X = torch.full((K, N), float(top), device="cuda")
    got = capi.bitMM2Int(da, capi.val2bit(X, b, True, False), M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05)
    assert got.shape == (M, N)
    assert bool((got == float(K * top)).all()), (float(got.min()), float(got.max()), K * top)
    # (2) random features: every row of the result == the exact column sums (odd and even low bits on top of large sums)
    g = torch.Generator(device="cuda").manual_seed(K % 9973 + N + b)
    Xr = torch.randint(0, top + 1, (K, N), device="cuda", generator=g)
    want = Xr.sum(dim=0, dtype=torch.int64)
    dx = capi.val2bit(Xr.float(), b, True, False)
    got = capi.bitMM2Int(da, dx, M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05)
    assert torch.equal(got.to(torch.int64), want.view(1, N).expand(M, N))
    # (3) the requantizing epilogue on the same accumulators agrees with the integer engine (exact int32)
    if K <= 65536:
        ob = 8
        bits_tc = capi.bitMM2Bit(da, dx, M, K, N, 1, b, ob, engine=capi.ENGINE_TCGEN05)
        bits_mma = capi.bitMM2Bit(da, dx, M, K, N, 1, b, ob, engine=capi.ENGINE_MMA_SYNC)
        assert torch.equal(bits_tc, bits_mma)


def test_tc05_f4_dense_16384_vs_c_oracle(capi, coracle):
    """BASELINE configs[4]'s largest adjacency, dense (Bernoulli 0.9 rows sum to ~44K at 2 bits), against the C oracle."""
    M = K = 16384
    N, b = 64, 2
    g = torch.Generator(device="cuda").manual_seed(11)
    A = (torch.rand((M, K), device="cuda", generator=g) < 0.9).float()
    X = torch.randint(0, 4, (K, N), device="cuda", generator=g).float()
    da, dx = capi.val2bit(A, 1, False, False), capi.val2bit(X, b, True, False)
    del A
    got = capi.bitMM2Int(da, dx, M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05).cpu().numpy()
    want = coracle.bitMM2Int(host_u32(da), host_u32(dx), M, K, N, 1, b, True)
    assert np.array_equal(got, want)
    gb = capi.bitMM2Bit(da, dx, M, K, N, 1, b, 8, engine=capi.ENGINE_TCGEN05)
    assert np.array_equal(host_u32(gb), coracle.bitMM2Bit(host_u32(da), host_u32(dx), M, K, N, 1, b, 8))


# BASELINE configs[1], the whole Fig. 8a grid with random inputs (2_7c_QGTC_GEMM_INT8.py shapes), tcgen05 engine
@pytest.mark.parametrize("b", [1, 2, 4, 8])
@pytest.mark.parametrize("N", [16, 32, 64])
@pytest.mark.parametrize("M", [1024, 2048, 4096])
def test_tc05_fig8a_grid_random(capi, coracle, M, N, b):
    K = M
    g = torch.Generator(device="cuda").manual_seed(M + 13 * N + b)
    A = (torch.rand((M, K), device="cuda", generator=g) < 0.5).float()
    X = torch.randint(0, 1 << b, (K, N), device="cuda", generator=g).float()
    da, dx = capi.val2bit(A, 1, False, False), capi.val2bit(X, b, True, False)
    ha, hx = host_u32(da), host_u32(dx)
    got = capi.bitMM2Bit(da, dx, M, K, N, 1, b, b, engine=capi.ENGINE_TCGEN05)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ha, hx, M, K, N, 1, b, b))
    gi = capi.bitMM2Int(da, dx, M, K, N, 1, b, True, engine=capi.ENGINE_TCGEN05)
    assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(ha, hx, M, K, N, 1, b, True))

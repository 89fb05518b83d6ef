"""GPU parity tests of the grouped tcgen05 engine (gemm_grp.cu; engine=ENGINE_GROUPED forced / qgtc_bitmm_grouped):
multi-plane A operands, in-kernel operand conversion, zero-extension on raw memory, all three epilogues, variable-shape
problem lists and whole driver chains issued op-major -- against the CPU oracle, bit-exact (float outputs: tolerance 0)."""
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


def operands(o, M, K, N, b1, b2, seed, density=0.3):
    rng = np.random.default_rng(seed)
    A = rand_adj(rng, M, K, density, zero_blocks=True) if b1 == 1 else rand_features(rng, (M, K), b1)
    X = rand_features(rng, (K, N), b2)
    return o.val2bit(A, b1, False, False), o.val2bit(X, b2, True, False), o.val2bit(X, b2, True, True)


# (M, K, N, bit1, bit2, out_bit): the GEMM cases of test_ops_gpu plus wide N, long K and every plane count of A
CASES = [
    (3, 3, 3, 1, 2, 2), (1, 1, 1, 1, 1, 1), (8, 128, 8, 1, 1, 1), (37, 300, 21, 1, 3, 3), (64, 128, 16, 1, 1, 1),
    (130, 513, 70, 2, 2, 2), (200, 1000, 121, 2, 2, 2), (255, 129, 65, 3, 5, 4), (129, 257, 33, 8, 8, 8),
    (759, 759, 16, 1, 2, 2), (759, 50, 16, 2, 2, 2), (16, 4096, 64, 1, 8, 8), (300, 300, 200, 1, 4, 4),
    (1000, 40, 40, 4, 4, 4), (33, 2000, 17, 1, 1, 8), (2258, 2258, 16, 1, 4, 4), (2258, 128, 16, 8, 8, 8),
    (515, 1283, 64, 1, 2, 2), (257, 640, 256, 1, 8, 8), (640, 700, 129, 5, 3, 6), (1283, 29, 64, 2, 2, 2),
    (4096, 4096, 64, 1, 2, 2), (384, 6000, 32, 7, 2, 3),
]


@pytest.mark.parametrize("M,K,N,b1,b2,ob", CASES)
def test_grouped_engine_single_problem_all_epilogues(capi, coracle, M, K, N, b1, b2, ob):
    ba, bx, bx8 = operands(coracle, M, K, N, b1, b2, seed=M + 3 * K + 7 * N + b1)
    da, dx, dx8 = dev_words(ba), dev_words(bx), dev_words(bx8)
    E = capi.ENGINE_GROUPED
    got = capi.bitMM2Bit(da, dx, M, K, N, b1, b2, ob, engine=E)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ba, bx, M, K, N, b1, b2, ob))
    assert torch.equal(capi.bitMM2Bit(da, dx, M, K, N, b1, b2, ob, engine=E, flags=capi.FLAG_NO_ZERO_JUMP), got)
    gi = capi.bitMM2Int(da, dx, M, K, N, b1, b2, True, engine=E)
    assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(ba, bx, M, K, N, b1, b2, True))
    g8 = capi.bitMM2Int(da, dx8, M, K, N, b1, b2, False, engine=E)
    assert np.array_equal(g8.cpu().numpy(), coracle.bitMM2Int(ba, bx8, M, K, N, b1, b2, False))
    gc = capi.bitMM2Bit_col(da, dx, M, K, N, b1, b2, ob, engine=E)
    want_c = coracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob)
    assert tuple(gc.shape) == want_c.shape and np.array_equal(host_u32(gc), want_c)


def test_grouped_engine_col_output_needs_no_prezeroed_buffer(capi, coracle):
    M, K, N, b = 300, 200, 21, 3
    ba, bx, _ = operands(coracle, M, K, N, b, b, seed=5)
    outs = [torch.full(capi.packed_shape(capi.LAYOUT_COL128, b, M, N), -1, dtype=torch.int32, device="cuda")]
    capi.bitmm_grouped([dev_words(ba)], [dev_words(bx)], [M], [K], N, b, b, b, kind=capi.GROUP_BITS_COL, outs=outs, engine=capi.ENGINE_GROUPED)
    assert np.array_equal(host_u32(outs[0]), coracle.bitMM2Bit_col(ba, bx, M, K, N, b, b, b))
    outs = [torch.full(capi.packed_shape(capi.LAYOUT_ROW, b, M, N), -1, dtype=torch.int32, device="cuda")]
    capi.bitmm_grouped([dev_words(ba)], [dev_words(bx)], [M], [K], N, b, b, b, outs=outs, engine=capi.ENGINE_GROUPED)
    assert np.array_equal(host_u32(outs[0]), coracle.bitMM2Bit(ba, bx, M, K, N, b, b, b))


@pytest.mark.parametrize("b1,cut1,cut2", [(1, 0, 37), (1, 501, 0), (2, 1000, 64), (2, 3, 1)])
def test_grouped_engine_zero_extension(capi, coracle, b1, cut1, cut2):
    """Operands that end before the reference's index formulas do (mismatched chains, SURVEY 8c rule 10): words beyond the
    buffer read as zero, also when the cut falls inside a 16-byte piece."""
    M, K, N, b2 = 300, 700, 40, 2
    ba, bx, _ = operands(coracle, M, K, N, b1, b2, seed=17 + cut1 + cut2)
    a = ba.reshape(-1)[: ba.size - cut1].copy()
    x = bx.reshape(-1)[: bx.size - cut2].copy()
    got = capi.bitMM2Bit(dev_words(a), dev_words(x), M, K, N, b1, b2, 4, engine=capi.ENGINE_GROUPED)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(a, x, M, K, N, b1, b2, 4))
    gi = capi.bitMM2Int(dev_words(a), dev_words(x), M, K, N, b1, b2, True, engine=capi.ENGINE_GROUPED)
    assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(a, x, M, K, N, b1, b2, True))


@pytest.mark.parametrize("N,b1,b2,ob,kind", [(16, 1, 2, 2, 0), (64, 1, 2, 2, 0), (16, 2, 2, 2, 0), (121, 4, 4, 4, 0), (16, 1, 8, 1, 1),
                                             (29, 1, 2, 2, 0), (40, 3, 3, 3, 2), (256, 1, 1, 1, 0)])
def test_grouped_variable_shapes(capi, coracle, N, b1, b2, ob, kind):
    """One launch, problems of different M and K (ragged tiles, a single-row problem, K shorter than one chunk)."""
    rng = np.random.default_rng(N + b1)
    shapes = [(759, 759), (128, 128), (1, 5), (300, 2000), (1283, 640), (130, 129), (8, 8), (2258, 300), (777, 1000)]
    if b1 > 1:
        shapes = [(m, min(k, 300)) for m, k in shapes]
    x1s, x2s, want = [], [], []
    for i, (M, K) in enumerate(shapes):
        ba, bx, bx8 = operands(coracle, M, K, N, b1, b2, seed=100 * i + N)
        x1s.append(ba)
        if kind == capi.GROUP_FLOAT:
            x2s.append(bx8)
            want.append(coracle.bitMM2Int(ba, bx8, M, K, N, b1, b2, False))
        elif kind == capi.GROUP_BITS_ROW:
            x2s.append(bx)
            want.append(coracle.bitMM2Bit(ba, bx, M, K, N, b1, b2, ob))
        else:
            x2s.append(bx)
            want.append(coracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob))
    Ms, Ks = [s[0] for s in shapes], [s[1] for s in shapes]
    for eng in (capi.ENGINE_GROUPED, capi.ENGINE_AUTO, capi.ENGINE_MMA_SYNC):
        outs = capi.bitmm_grouped([dev_words(a) for a in x1s], [dev_words(x) for x in x2s], Ms, Ks, N, b1, b2, ob, kind=kind,
                                  pad_128=False, engine=eng)
        for i, (g, w) in enumerate(zip(outs, want)):
            gh = g.cpu().numpy()
            assert np.array_equal(gh.view(np.uint32) if gh.dtype == np.int32 else gh, w), (eng, i, shapes[i])


def test_grouped_with_tile_bitmaps(capi, oracle, coracle):
    """Zero-tile jumping from per-problem tile bitmaps: block-diagonal adjacencies (batched-GIN shaped), an empty graph, a
    dense one; bit-identical to the oracle and to the run without bitmaps."""
    N, b = 64, 2
    rng = np.random.default_rng(9)
    x1s, x2s, bms, Ms = [], [], [], []
    for i, n in enumerate([1283, 640, 2000, 300, 1500]):
        A = np.zeros((n, n), np.float32)
        if i != 3:
            for b0 in range(0, n, 40):                      # small graphs of ~40 nodes on the diagonal
                e = min(n, b0 + 40)
                A[b0:e, b0:e] = rng.random((e - b0, e - b0)) < 0.1
        if i == 4:
            A = (rng.random((n, n)) < 0.3).astype(np.float32)
        ba = coracle.val2bit(A, 1, False, False)
        x1s.append(ba)
        x2s.append(coracle.val2bit(rand_features(rng, (n, N), b), b, True, False))
        Ms.append(n)
    d1, d2 = [dev_words(a) for a in x1s], [dev_words(x) for x in x2s]
    bms = [capi.tile_bitmap(a, n, n) for a, n in zip(d1, Ms)]
    outs = capi.bitmm_grouped(d1, d2, Ms, Ms, N, 1, b, b, tile_maps=bms, engine=capi.ENGINE_GROUPED)
    plain = capi.bitmm_grouped(d1, d2, Ms, Ms, N, 1, b, b, engine=capi.ENGINE_GROUPED)
    for i, n in enumerate(Ms):
        assert np.array_equal(host_u32(outs[i]), coracle.bitMM2Bit(x1s[i], x2s[i], n, n, N, 1, b, b)), i
        assert torch.equal(outs[i], plain[i])
    fl = capi.bitmm_grouped(d1, d2, Ms, Ms, N, 1, b, 1, kind=capi.GROUP_FLOAT, pad_128=True, tile_maps=bms, engine=capi.ENGINE_GROUPED)
    for i, n in enumerate(Ms):
        assert np.array_equal(fl[i].cpu().numpy(), coracle.bitMM2Int(x1s[i], x2s[i], n, n, N, 1, b, True)), i


def test_grouped_many_problems_and_repeats(capi, coracle):
    """More problems than one launch's table holds (split into several launches), tiny ragged problems, repeated: the
    pipeline items alternate between one-chunk and multi-chunk stages across hundreds of units per CTA."""
    N, b = 16, 2
    rng = np.random.default_rng(21)
    pool = []
    for _ in range(40):
        M, K = int(rng.integers(1, 400)), int(rng.integers(1, 700))
        ba, bx, _ = operands(coracle, M, K, N, 1, b, seed=M * 1000 + K)
        pool.append((M, K, dev_words(ba), dev_words(bx), dev_words(coracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b))))
    picks = [pool[int(i)] for i in rng.integers(0, len(pool), size=500)]
    Ms, Ks = [u[0] for u in picks], [u[1] for u in picks]
    for _ in range(3):
        outs = capi.bitmm_grouped([u[2] for u in picks], [u[3] for u in picks], Ms, Ks, N, 1, b, b, engine=capi.ENGINE_GROUPED)
        torch.cuda.synchronize()
        bad = [i for i, (o, u) in enumerate(zip(outs, picks)) if not torch.equal(o, u[4])]
        assert not bad, bad[:10]


@pytest.mark.parametrize("count,M", [(24, 130), (24, 2700), (25, 130), (80, 300), (80, 2500), (81, 130), (224, 40)])
def test_grouped_parameter_block_sizes(capi, coracle, count, M):
    """The problem / unit tables travel as kernel parameters in the smallest of three block sizes (24 problems / 512 units,
    80 / 1536, 224 / 4096): launches on both sides of every boundary, by problem count and by unit count."""
    N, b, K = 16, 2, 200
    pool = []
    for i in range(4):
        m = M - 3 * i
        ba, bx, _ = operands(coracle, m, K, N, 1, b, seed=7000 + m)
        pool.append((m, dev_words(ba), dev_words(bx), dev_words(coracle.bitMM2Bit(ba, bx, m, K, N, 1, b, b))))
    picks = [pool[i % 4] for i in range(count)]
    outs = capi.bitmm_grouped([u[1] for u in picks], [u[2] for u in picks], [u[0] for u in picks], [K] * count, N, 1, b, b,
                              engine=capi.ENGINE_GROUPED)
    torch.cuda.synchronize()
    bad = [i for i, (o, u) in enumerate(zip(outs, picks)) if not torch.equal(o, u[3])]
    assert not bad, bad[:10]


@pytest.mark.parametrize("b,K,N", [(2, 50, 16), (4, 128, 121), (1, 300, 64), (8, 16, 16)])
def test_grouped_shared_weights_stage_reuse(capi, coracle, b, K, N):
    """The weight products of an epoch (X.W, o.W: every problem multiplies with the same packed W, K of one or a few chunks):
    the converters re-use a stage that already holds the converted chunk.  Hundreds of units per launch so that every CTA
    runs through several of them; a second weight matrix is interleaved so that runs of identical chunks start and stop."""
    rng = np.random.default_rng(b * 100 + K)
    Wa = coracle.val2bit(rand_features(rng, (K, N), b), b, True, False)
    Wb = coracle.val2bit(rand_features(rng, (K, N), b), b, True, False)
    dWa, dWb = dev_words(Wa), dev_words(Wb)
    pool = []
    for M in (759, 128, 300, 1, 515):
        x = coracle.val2bit(rand_features(rng, (M, K), b), b, False, False)
        pool.append((M, dev_words(x), {0: dev_words(coracle.bitMM2Bit(x, Wa, M, K, N, b, b, b)), 1: dev_words(coracle.bitMM2Bit(x, Wb, M, K, N, b, b, b))}))
    picks = [(pool[int(i)], int(w)) for i, w in zip(rng.integers(0, len(pool), size=220), (rng.random(220) < 0.15))]
    outs = capi.bitmm_grouped([p[0][1] for p in picks], [dWb if p[1] else dWa for p in picks], [p[0][0] for p in picks], [K] * len(picks),
                              N, b, b, b, engine=capi.ENGINE_GROUPED)
    torch.cuda.synchronize()
    bad = [i for i, (o, p) in enumerate(zip(outs, picks)) if not torch.equal(o, p[0][2][p[1]])]
    assert not bad, bad[:10]


def _chain_specs(kind, n, D, h, C, b):
    """(A operand, B operand, K, N, bit1, float?) per op; names refer to the running values of the chain."""
    if kind == "gcn":       # main_qgtc.py:146-154
        return [("X", "W1", D, h, b, False), ("A", "o", n, h, 1, False), ("o", "W2", h, h, b, False), ("A", "o", n, h, 1, False),
                ("o", "W3", h, C, b, False), ("A", "o", n, h, 1, True)]
    return [("A", "X", n, D, 1, False), ("o", "W1", D, h, b, False), ("A", "o", n, h, 1, False), ("o", "W2", h, h, b, False),
            ("A", "o", n, h, 1, False), ("o", "W3", h, C, b, True)]          # main_qgtc.py:131-138


@pytest.mark.parametrize("kind", ["gcn", "gin"])
@pytest.mark.parametrize("D,h,C,b", [(50, 16, 121, 2), (128, 16, 40, 4), (29, 64, 2, 2), (50, 16, 10, 8), (32, 16, 7, 1)])
def test_op_major_chains_match_oracle(capi, coracle, kind, D, h, C, b):
    """The forward of several batches issued op-major -- one grouped launch per op over all batches, which is how the
    drivers run an epoch -- equals the oracle's batch-major chains on raw memory (layout mismatches of the reference kept)."""
    rng = np.random.default_rng(D + h + b)
    ns = [759, 300, 1283, 70, 515]
    o = coracle
    W = {"W1": o.val2bit(np.ones((D, h), np.float32), b, True, False), "W2": o.val2bit(np.ones((h, h), np.float32), b, True, False),
         "W3": o.val2bit(np.ones((h, C), np.float32), b, True, True)}
    batches = []
    for n in ns:
        A = rand_adj(rng, n, n, 0.02)
        X = rng.standard_normal((n, D)).astype(np.float32)
        batches.append({"n": n, "A": o.val2bit(A, 1, False, False), "X": o.val2bit(X, b, True, False)})
    # oracle: batch-major
    want = []
    for bt in batches:
        vals = dict(W, A=bt["A"], X=bt["X"])
        for (a, x, K, N, b1, is_float) in _chain_specs(kind, bt["n"], D, h, C, b):
            vals["o"] = (o.bitMM2Int(vals[a], vals[x], bt["n"], K, N, b1, b, False) if is_float
                         else o.bitMM2Bit(vals[a], vals[x], bt["n"], K, N, b1, b, b))
        want.append(vals["o"])
    # device: op-major
    dW = {k: dev_words(v) for k, v in W.items()}
    vals = [dict(dW, A=dev_words(bt["A"]), X=dev_words(bt["X"])) for bt in batches]
    n_ops = 6
    for i in range(n_ops):
        specs = [_chain_specs(kind, bt["n"], D, h, C, b)[i] for bt in batches]
        a, x, _, N, b1, is_float = specs[0]
        outs = capi.bitmm_grouped([v[a] for v in vals], [v[x] for v in vals], ns, [s[2] for s in specs], N, b1, b, b,
                                  kind=capi.GROUP_FLOAT if is_float else capi.GROUP_BITS_ROW, pad_128=False, engine=capi.ENGINE_GROUPED)
        for v, out in zip(vals, outs):
            v["o"] = out
    for i, (v, w) in enumerate(zip(vals, want)):
        assert np.array_equal(v["o"].cpu().numpy(), w), i


@pytest.mark.parametrize("N,N2,b,kind2,pad", [(16, 16, 2, 0, False), (16, 121, 2, 0, False), (64, 64, 2, 0, False), (29, 64, 2, 0, False), (64, 2, 2, 1, False),
                                                (16, 16, 1, 0, False), (16, 40, 4, 0, False), (16, 16, 8, 0, False), (48, 128, 3, 1, True), (33, 7, 5, 1, False)])
def test_fused_aggregation_update_equals_two_ops(capi, coracle, N, N2, b, kind2, pad):
    """qgtc_bitmm_grouped_fused: out2 = op2(bitMM2Bit(A, X), W) in one launch == the two ops of main_qgtc.py:148+150 etc. through the
    oracle, for several batches of different size (with tile bitmaps, with and without writing the first op's own result)."""
    rng = np.random.default_rng(N * 1000 + N2 * 10 + b)
    ns = [759, 300, 1283, 70, 128, 1]
    W = coracle.val2bit(rand_features(rng, (N, N2), b), b, True, kind2 == 1 and not pad)
    x1s, x2s, want1, want2 = [], [], [], []
    for n in ns:
        A = rand_adj(rng, n, n, 0.03, zero_blocks=True)
        X = rand_features(rng, (n, N), b)
        ba, bx = coracle.val2bit(A, 1, False, False), coracle.val2bit(X, b, True, False)
        o1 = coracle.bitMM2Bit(ba, bx, n, n, N, 1, b, b)
        x1s.append(dev_words(ba)); x2s.append(dev_words(bx)); want1.append(o1)
        want2.append(coracle.bitMM2Int(o1, W, n, N, N2, b, b, pad) if kind2 == 1 else coracle.bitMM2Bit(o1, W, n, N, N2, b, b, b))
    dW = dev_words(W)
    maps = [capi.tile_bitmap(a, n, n).view(-1) for a, n in zip(x1s, ns)]
    for tm, w1 in ((None, False), (maps, True)):
        outs2, outs1 = capi.bitmm_grouped_fused(x1s, x2s, ns, ns, N, 1, b, b, dW, N2, b, b, kind2=kind2, pad_128=pad, tile_maps=tm, write_out1=w1,
                                                engine=capi.ENGINE_GROUPED)
        for i in range(len(ns)):
            g = outs2[i].cpu().numpy()
            assert np.array_equal(g.view(np.uint32) if g.dtype == np.int32 else g, want2[i]), (i, ns[i], w1)
            if w1:
                assert np.array_equal(host_u32(outs1[i]), want1[i]), (i, ns[i])


def test_fused_many_units_per_cta(capi, coracle):
    """Hundreds of units per launch: every CTA runs several fused units back to back (A2 / D2 re-use across units)."""
    N, N2, b = 16, 16, 2
    rng = np.random.default_rng(5)
    W = coracle.val2bit(rand_features(rng, (N, N2), b), b, True, False)
    pool = []
    for n in (759, 515, 300, 129):
        ba = coracle.val2bit(rand_adj(rng, n, n, 0.02), 1, False, False)
        bx = coracle.val2bit(rand_features(rng, (n, N), b), b, True, False)
        o1 = coracle.bitMM2Bit(ba, bx, n, n, N, 1, b, b)
        pool.append((n, dev_words(ba), dev_words(bx), dev_words(coracle.bitMM2Bit(o1, W, n, N, N2, b, b, b))))
    picks = [pool[int(i)] for i in rng.integers(0, len(pool), size=150)]
    ns = [p[0] for p in picks]
    for _ in range(2):
        outs2, _ = capi.bitmm_grouped_fused([p[1] for p in picks], [p[2] for p in picks], ns, ns, N, 1, b, b, dev_words(W), N2, b, b,
                                            engine=capi.ENGINE_GROUPED)
        torch.cuda.synchronize()
        bad = [i for i, (o, p) in enumerate(zip(outs2, picks)) if not torch.equal(o, p[3])]
        assert not bad, bad[:10]


def test_fused_refuses_shapes_outside_the_kernel(capi):
    a = torch.zeros((8, 4), dtype=torch.int32, device="cuda")
    x = torch.zeros((4, 128), dtype=torch.int32, device="cuda")
    w = torch.zeros((4, 128), dtype=torch.int32, device="cuda")
    with pytest.raises(capi.QgtcError, match="unsupported"):
        capi.bitmm_grouped_fused([a], [x], [8], [8], 100, 1, 1, 1, w, 8, 1, 1)       # N = 100 > 64: the caller issues two launches


def test_grouped_rejects_bad_arguments(capi):
    a = torch.zeros((8, 4), dtype=torch.int32, device="cuda")
    x = torch.zeros((4, 128), dtype=torch.int32, device="cuda")
    small = [torch.zeros((3,), dtype=torch.int32, device="cuda")]
    with pytest.raises(capi.QgtcError, match="too small"):
        capi.bitmm_grouped([a], [x], [8], [8], 8, 1, 1, 1, outs=small)
    with pytest.raises(capi.QgtcError, match="unsupported"):
        capi.bitmm_grouped([a], [x], [8], [8], 8, 9, 1, 1)
    with pytest.raises(capi.QgtcError):
        capi.bitmm_grouped([a], [x], [8], [8], 300, 1, 1, 1, engine=capi.ENGINE_GROUPED)   # N > 256: not on this engine

"""GPU parity tests: the CUDA path, called through the C ABI (qgtc_ppopp22_b200.capi) and through
the `QGTC` extension, against the CPU oracle on the same seeded inputs.  Integer results are
compared bit-exactly; the only floating-point output (bitMM2Int) is an exact int32->fp32
conversion and is compared exactly too (tolerance 0)."""
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


@pytest.fixture(scope="module")
def QGTC():
    import qgtc_ppopp22_b200
    return qgtc_ppopp22_b200.load_extension()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def dev_words(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host_u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


# ------------------------------------------------------------------------------ val2bit / bit2val

PACK_SHAPES = [(3, 3), (8, 128), (37, 300), (130, 513), (759, 50), (64, 121), (1, 1), (257, 4100), (16, 10), (500, 16)]


@pytest.mark.parametrize("H,W", PACK_SHAPES)
@pytest.mark.parametrize("nbits", [1, 2, 3, 4, 8])
def test_val2bit_all_layouts(capi, oracle, H, W, nbits):
    rng = np.random.default_rng(H * 131 + W * 7 + nbits)
    x = rand_features(rng, (H, W), nbits)
    x[0, 0] = np.nan
    xd = dev(x)
    for cm, ol in [(False, False), (True, False), (True, True)]:
        got = capi.val2bit(xd, nbits, cm, ol)
        want = oracle.val2bit(x, nbits, cm, ol)
        assert tuple(got.shape) == want.shape and got.dtype == torch.int32
        assert np.array_equal(host_u32(got), want), (cm, ol)
        # round trip through bit2val (values masked to nbits, q == 2^b -> 0)
        back = capi.bit2val(got, nbits, H, W, cm, ol)
        if not (cm and ol):
            assert np.array_equal(back.cpu().numpy(), oracle.quantize_val(x, nbits) & ((1 << nbits) - 1))
        assert np.array_equal(back.cpu().numpy(), oracle.bit2val(want, nbits, H, W, cm, ol))


@pytest.mark.parametrize("nbits", [1, 2, 8])
def test_val2bit_batched_equals_per_matrix(capi, oracle, nbits):
    """qgtc_val2bit_batched: a list of ragged matrices (odd widths: rows only 4-byte aligned; offset storage; tiny and empty-ish
    shapes) packed in one call equals QGTC.val2bit of every matrix, in all three layouts, and the oracle for the ROW layout."""
    rng = np.random.default_rng(91 + nbits)
    shapes = [(759, 759), (130, 129), (1, 5), (8, 128), (300, 2001), (64, 64), (17, 1023), (515, 33), (2, 3)]
    xs = [rand_features(rng, sh, nbits) for sh in shapes]
    xs[0][0, 0] = np.nan
    # one matrix lives at an odd offset inside a larger buffer (its first row is not 16-byte aligned)
    big = torch.zeros(3 + 130 * 129, dtype=torch.float32, device="cuda")
    big[3:] = dev(xs[1]).reshape(-1)
    xd = [dev(x) for x in xs]
    xd[1] = big[3:].view(130, 129)
    for cm, ol in [(False, False), (True, False), (True, True)]:
        got = capi.val2bit_batched(xd, nbits, cm, ol)
        assert len(got) == len(xs)
        for x, xdev, g in zip(xs, xd, got):
            one = capi.val2bit(xdev, nbits, cm, ol)
            assert g.shape == one.shape and torch.equal(g, one), (x.shape, cm, ol)
            if not cm:
                assert np.array_equal(host_u32(g), oracle.val2bit(x, nbits, cm, ol)), x.shape
    # a list of wide matrices only (every row >= 16 output words): the word-per-lane form, rows misaligned by 0..3 floats
    wide_shapes = [(759, 759), (300, 2001), (17, 1023), (515, 513), (64, 512), (9, 4096), (130, 601), (33, 1022)]
    ws = [rand_features(rng, sh, nbits) for sh in wide_shapes]
    wd = [dev(x) for x in ws]
    for off, i in ((1, 6), (2, 3), (3, 0)):      # storage offsets of 1, 2, 3 floats
        h, w = wide_shapes[i]
        buf = torch.zeros(off + h * w, dtype=torch.float32, device="cuda")
        buf[off:] = wd[i].reshape(-1)
        wd[i] = buf[off:].view(h, w)
    got = capi.val2bit_batched(wd, nbits, False, False)
    for x, xdev, g in zip(ws, wd, got):
        assert torch.equal(g, capi.val2bit(xdev, nbits, False, False)), x.shape
        assert np.array_equal(host_u32(g), oracle.val2bit(x, nbits, False, False)), x.shape
    # the same entry point through the QGTC extension
    import qgtc_ppopp22_b200
    Q = qgtc_ppopp22_b200.load_extension()
    for g, e in zip(got, Q.val2bit_batched(wd, nbits, False, False)):
        assert torch.equal(g, e)
    # a list longer than one launch's problem table (448)
    many = [dev(rand_features(rng, (9 + (i % 5), 40 + (i % 7)), nbits)) for i in range(500)]
    got = capi.val2bit_batched(many, nbits, False, False)
    for x, g in zip(many, got):
        assert torch.equal(g, capi.val2bit(x, nbits, False, False))


def test_val2bit_full_size_adjacency(capi, coracle):
    # BASELINE config[1]: 4096 x 4096 1-bit adjacency (64 MiB fp32 -> 2 MiB packed)
    rng = np.random.default_rng(3)
    a = rand_adj(rng, 4096, 4096, 0.3)
    got = capi.val2bit(dev(a), 1, False, False)
    assert np.array_equal(host_u32(got), coracle.val2bit(a, 1, False, False))
    # size-independent property: popcount of the packed tensor == number of ones
    assert int(np.unpackbits(host_u32(got).view(np.uint8)).sum()) == int(a.sum())
    x = rand_features(rng, (4096, 64), 8)
    assert np.array_equal(host_u32(capi.val2bit(dev(x), 8, True, False)), coracle.val2bit(x, 8, True, False))


def test_val2bit_unaligned_view(capi, oracle):
    # odd widths / offset storage take the scalar load path
    rng = np.random.default_rng(5)
    big = rand_features(rng, (65, 131), 3)
    t = dev(big)[1:, :].contiguous()
    want = oracle.val2bit(big[1:, :], 3, False, False)
    assert np.array_equal(host_u32(capi.val2bit(t, 3, False, False)), want)
    flat = dev(np.concatenate([np.zeros(1, np.float32), big.reshape(-1)]))
    view = flat[1:].view(65, 131)          # data_ptr is 4-byte aligned only
    assert np.array_equal(host_u32(capi.val2bit(view, 3, False, False)), oracle.val2bit(big, 3, False, False))


# ------------------------------------------------------------------------------ bit GEMMs

GEMM_CASES = [
    (3, 3, 3, 1, 2, 2), (1, 1, 1, 1, 1, 1), (8, 128, 8, 1, 1, 1), (37, 300, 21, 1, 3, 3), (64, 128, 16, 1, 1, 1),
    (130, 513, 70, 2, 2, 2), (200, 1000, 121, 2, 2, 2), (255, 129, 65, 3, 5, 4), (129, 257, 33, 8, 8, 8),
    (759, 759, 16, 1, 2, 2), (759, 50, 16, 2, 2, 2), (16, 4096, 64, 1, 8, 8), (300, 300, 200, 1, 4, 4),
    (1000, 40, 40, 4, 4, 4), (33, 2000, 17, 1, 1, 8),
]
ENGINES = [0, 1]  # AUTO, MMA_SYNC  (TCGEN05 has its own shape-restricted tests)


def make_operands(oracle, M, K, N, b1, b2, seed, density=0.3):
    rng = np.random.default_rng(seed)
    A = rand_adj(rng, M, K, density, zero_blocks=True) if b1 == 1 else rand_features(rng, (M, K), b1)
    X = rand_features(rng, (K, N), b2)
    return A, X, oracle.val2bit(A, b1, False, False), oracle.val2bit(X, b2, True, False), oracle.val2bit(X, b2, True, True)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,K,N,b1,b2,ob", GEMM_CASES)
def test_bitmm2bit(capi, oracle, M, K, N, b1, b2, ob, engine):
    A, X, ba, bx, _ = make_operands(oracle, M, K, N, b1, b2, seed=M + 3 * K + 5 * N)
    got = capi.bitMM2Bit(dev_words(ba), dev_words(bx), M, K, N, b1, b2, ob, engine=engine)
    want = oracle.bitMM2Bit(ba, bx, M, K, N, b1, b2, ob)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(host_u32(got), want)
    # zero-tile jumping never changes the result
    nj = capi.bitMM2Bit(dev_words(ba), dev_words(bx), M, K, N, b1, b2, ob, engine=engine, flags=capi.FLAG_NO_ZERO_JUMP)
    assert torch.equal(nj, got)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,K,N,b1,b2,ob", GEMM_CASES)
def test_bitmm2int(capi, oracle, M, K, N, b1, b2, ob, engine):
    A, X, ba, bx, bx8 = make_operands(oracle, M, K, N, b1, b2, seed=2 * M + K + N)
    got = capi.bitMM2Int(dev_words(ba), dev_words(bx), M, K, N, b1, b2, True, engine=engine)
    want = oracle.bitMM2Int(ba, bx, M, K, N, b1, b2, True)
    assert got.dtype == torch.float32 and tuple(got.shape) == (M, N)
    assert np.array_equal(got.cpu().numpy(), want)          # exact: tolerance 0
    got8 = capi.bitMM2Int(dev_words(ba), dev_words(bx8), M, K, N, b1, b2, False, engine=engine)
    assert np.array_equal(got8.cpu().numpy(), oracle.bitMM2Int(ba, bx8, M, K, N, b1, b2, False))


@pytest.mark.parametrize("M,K,N,b1,b2,ob", GEMM_CASES)
def test_bitmm2bit_col(capi, oracle, M, K, N, b1, b2, ob):
    A, X, ba, bx, _ = make_operands(oracle, M, K, N, b1, b2, seed=M * K + N)
    got = capi.bitMM2Bit_col(dev_words(ba), dev_words(bx), M, K, N, b1, b2, ob)
    want = oracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(host_u32(got), want)
    quirk = capi.bitMM2Bit_col(dev_words(ba), dev_words(bx), M, K, N, b1, b2, ob, flags=capi.FLAG_COL_REF_QUIRK)
    assert np.array_equal(host_u32(quirk), oracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob, ref_quirk=True))


@pytest.mark.parametrize("b", [1, 2, 4, 8])
@pytest.mark.parametrize("engine", ENGINES)
def test_headline_shape_vs_c_oracle(capi, coracle, b, engine):
    # BASELINE config[1]: 4096 x 4096 x 64, 1-bit adjacency x b-bit features, requantized to b bits
    M = K = 4096
    N = 64
    rng = np.random.default_rng(b)
    A, X = rand_adj(rng, M, K, 0.01 if b == 2 else 0.5, zero_blocks=True), rand_features(rng, (K, N), b)
    ba, bx = coracle.val2bit(A, 1, False, False), coracle.val2bit(X, b, True, False)
    got = capi.bitMM2Bit(dev_words(ba), dev_words(bx), M, K, N, 1, b, b, engine=engine)
    assert np.array_equal(host_u32(got), coracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b))
    gi = capi.bitMM2Int(dev_words(ba), dev_words(bx), M, K, N, 1, b, True, engine=engine)
    assert np.array_equal(gi.cpu().numpy(), coracle.bitMM2Int(ba, bx, M, K, N, 1, b, True))
    # size-independent property: equals the value-level integer product (float64 is exact here)
    xq = torch.from_numpy((np.rint(np.clip(np.where(X < 0, 1, np.where(X > 2 ** b, 2 ** b - 1, X)), 0, None)).astype(np.int64)
                           & (2 ** b - 1)).astype(np.float64)).cuda()
    ref = (torch.from_numpy(A.astype(np.float64)).cuda() @ xq).to(torch.float32)
    assert torch.equal(gi, ref)


@pytest.mark.parametrize("M,K,N", [(1024, 1024, 16), (2048, 2048, 32), (4096, 4096, 16), (2258, 2258, 128)])
def test_sweep_shapes_all_ones(capi, M, K, N):
    # reference throughput inputs (2_7c_QGTC_GEMM_INT8.py:6-7): all-ones => bitMM2Int == K, bitMM2Bit == sat(K)
    ba = capi.val2bit(torch.ones(M, K, device="cuda"), 1, False, False)
    for b in (1, 2, 4, 8):
        bx = capi.val2bit(torch.ones(K, N, device="cuda"), b, True, False)
        assert bool((capi.bitMM2Int(ba, bx, M, K, N, 1, b, True) == K).all())
        out = capi.bit2val(capi.bitMM2Bit(ba, bx, M, K, N, 1, b, b), b, M, N, False)
        assert bool((out == (2 ** b - 1)).all())


def gcn_chain(ops, bit_A, bit_X, bW1, bW2, bW3, n, D, h, C, b):
    """main_qgtc.py:146-154, including its layout mismatches (ops on raw memory)."""
    o = ops.bitMM2Bit(bit_X, bW1, n, D, h, b, b, b)
    o1 = ops.bitMM2Bit(bit_A, o, n, n, h, 1, b, b)
    o2 = ops.bitMM2Bit(o1, bW2, n, h, h, b, b, b)
    o3 = ops.bitMM2Bit(bit_A, o2, n, n, h, 1, b, b)
    o4 = ops.bitMM2Bit(o3, bW3, n, h, C, b, b, b)
    return [o, o1, o2, o3, o4, ops.bitMM2Int(bit_A, o4, n, n, h, 1, b, False)]


def gin_chain(ops, bit_A, bit_X, bW1, bW2, bW3, n, D, h, C, b):
    """main_qgtc.py:131-138."""
    o = ops.bitMM2Bit(bit_A, bit_X, n, n, D, 1, b, b)
    o1 = ops.bitMM2Bit(o, bW1, n, D, h, b, b, b)
    o2 = ops.bitMM2Bit(bit_A, o1, n, n, h, 1, b, b)
    o3 = ops.bitMM2Bit(o2, bW2, n, h, h, b, b, b)
    o4 = ops.bitMM2Bit(bit_A, o3, n, n, h, 1, b, b)
    return [o, o1, o2, o3, o4, ops.bitMM2Int(o4, bW3, n, h, C, b, b, False)]


@pytest.mark.parametrize("chain", [gcn_chain, gin_chain])
@pytest.mark.parametrize("n,D,h,C,b", [(759, 50, 16, 121, 2), (300, 128, 16, 40, 1), (515, 29, 64, 2, 2), (70, 50, 16, 10, 4)])
def test_driver_chains_on_raw_memory(capi, coracle, chain, n, D, h, C, b):
    """The reference drivers feed row-major outputs into the column-major slot; every op is
    defined on raw memory with zero-extension, so whole chains must agree word for word."""
    rng = np.random.default_rng(n + D)
    A = rand_adj(rng, n, n, 0.02)
    X = rng.standard_normal((n, D)).astype(np.float32)          # dataset.py:64 uses torch.randn features
    W1, W2, W3 = np.ones((D, h), np.float32), np.ones((h, h), np.float32), np.ones((h, C), np.float32)
    o = coracle
    pk = dict(bit_A=o.val2bit(A, 1, False, False), bit_X=o.val2bit(X, b, True, False), bW1=o.val2bit(W1, b, True, False),
              bW2=o.val2bit(W2, b, True, False), bW3=o.val2bit(W3, b, True, True))
    want = chain(o, n=n, D=D, h=h, C=C, b=b, **pk)
    got = chain(capi, n=n, D=D, h=h, C=C, b=b, **{k: dev_words(v) for k, v in pk.items()})
    for i, (g, w) in enumerate(zip(got, want)):
        gh = g.cpu().numpy()
        assert np.array_equal(gh.view(np.uint32) if gh.dtype == np.int32 else gh, w), "op %d of the chain differs" % (i + 1)


def test_batched_and_host_entry_points(capi, coracle):
    M, K, N, b = 256, 384, 64, 2
    rng = np.random.default_rng(11)
    cnt = 5
    x1 = np.stack([coracle.val2bit(rand_adj(rng, M, K, 0.1), 1, False, False).reshape(-1) for _ in range(cnt)])
    x2 = np.stack([coracle.val2bit(rand_features(rng, (K, N), b), b, True, False).reshape(-1) for _ in range(cnt)])
    out = capi.bitMM2Bit_batched(dev_words(x1), dev_words(x2), cnt, M, K, N, 1, b, b)
    for i in range(cnt):
        assert np.array_equal(host_u32(out[i]).reshape(-1), coracle.bitMM2Bit(x1[i], x2[i], M, K, N, 1, b, b).reshape(-1))
    # host-buffer form (pinned in, pinned out)
    h1 = torch.from_numpy(x1[0].view(np.int32)).pin_memory()
    h2 = torch.from_numpy(x2[0].view(np.int32)).pin_memory()
    ho = capi.bitMM2Bit_host(h1, h2, M, K, N, 1, b, b)
    assert np.array_equal(ho.numpy().view(np.uint32).reshape(-1), coracle.bitMM2Bit(x1[0], x2[0], M, K, N, 1, b, b).reshape(-1))
    assert capi.launch_count() > 0
    # pipelined host-batched form: more problems than one group holds, ragged last group
    cnt2 = 11
    hx1 = torch.from_numpy(np.stack([x1[i % cnt] for i in range(cnt2)]).view(np.int32)).pin_memory()
    hx2 = torch.from_numpy(np.stack([x2[i % cnt] for i in range(cnt2)]).view(np.int32)).pin_memory()
    import os
    os.environ["QGTC_HOST_GROUP"] = "4"            # 3 groups, the last one ragged
    try:
        hob = capi.bitMM2Bit_host_batched(hx1, hx2, cnt2, M, K, N, 1, b, b)
    finally:
        del os.environ["QGTC_HOST_GROUP"]
    for i in range(cnt2):
        assert np.array_equal(hob[i].numpy().view(np.uint32), coracle.bitMM2Bit(x1[i % cnt], x2[i % cnt], M, K, N, 1, b, b).reshape(-1)), i


def test_tile_counts(capi, oracle):
    rng = np.random.default_rng(2)
    M, K, N = 300, 1000, 40
    A = rand_adj(rng, M, K, 0.002)
    ba = oracle.val2bit(A, 1, False, False)
    assert capi.tile_counts(dev_words(ba), M, K, N, 1, 2) == oracle.tile_counts(ba, M, K, N, 1, 2)
    X = rand_features(rng, (M, K), 3)
    bx = oracle.val2bit(X, 3, False, False)
    assert capi.tile_counts(dev_words(bx), M, K, N, 3, 2) == oracle.tile_counts(bx, M, K, N, 3, 2)


# ------------------------------------------------------------------------------ the `QGTC` extension

def test_extension_ops_match_capi_and_oracle(QGTC, capi, oracle, capfd):
    M, K, N, b = 130, 513, 70, 2
    A, X, ba, bx, bx8 = make_operands(oracle, M, K, N, 1, b, seed=9)
    Ad, Xd = dev(A), dev(X)
    bit_a = QGTC.val2bit(Ad, 1, False, False)
    bit_x = QGTC.val2bit(Xd, b, True, False)
    bit_x8 = QGTC.val2bit(Xd, b, True, True)
    assert np.array_equal(host_u32(bit_a), ba) and np.array_equal(host_u32(bit_x), bx) and np.array_equal(host_u32(bit_x8), bx8)
    assert bit_a.dtype == torch.int32 and tuple(bit_x.shape) == (b * 4 * ((K + 127) // 128), 128)
    out = QGTC.bitMM2Bit(bit_a, bit_x, M, K, N, 1, b, b)
    assert np.array_equal(host_u32(out), oracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b))
    fo = QGTC.bitMM2Int(bit_a, bit_x8, M, K, N, 1, b, False)
    assert np.array_equal(fo.cpu().numpy(), oracle.bitMM2Int(ba, bx8, M, K, N, 1, b, False))
    col = QGTC.bitMM2Bit_col(bit_a, bit_x, M, K, N, 1, b, b)
    assert np.array_equal(host_u32(col), oracle.bitMM2Bit_col(ba, bx, M, K, N, 1, b, b))
    val = QGTC.bit2val(out, b, M, N, False, False)
    assert np.array_equal(val.cpu().numpy(), oracle.bit2val(oracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b), b, M, N, False))
    capfd.readouterr()
    # counters: cumulative, printed like the reference (parse_counter.py greps these lines)
    t, z = oracle.tile_counts(ba, M, K, N, 1, b)
    o1 = QGTC.bitMM2Bit_base_cnt(bit_a, bit_x, M, K, N, 1, b, b)
    o2 = QGTC.bitMM2Bit_zerojump_cnt(bit_a, bit_x, M, K, N, 1, b, b)
    txt = capfd.readouterr().out
    assert "counter_global: %d" % t in txt and "counter: %d" % z in txt
    assert torch.equal(o1, out) and torch.equal(o2, out)
    o3 = QGTC.bitMM2Bit_profile(bit_a, bit_x, M, K, N, 1, b, b)
    txt = capfd.readouterr().out
    assert "X1_height %d, X1_width: %d, X2_width: %d, TFLOPs:" % (M, K, N) in txt
    assert torch.equal(o3, out)


def test_extension_error_behaviour(QGTC):
    with pytest.raises(RuntimeError, match="must be contiguous"):
        QGTC.val2bit(torch.ones(8, 8, device="cuda").t(), 2, False, False)
    with pytest.raises(RuntimeError):
        QGTC.val2bit(torch.ones(8, 8, device="cuda", dtype=torch.float64), 2, False, False)
    with pytest.raises(RuntimeError):
        QGTC.bitMM2Bit(torch.zeros(8, 4, device="cuda"), torch.zeros(4, 128, device="cuda", dtype=torch.int32), 8, 8, 8, 1, 1, 1)
    # the process is still alive and the device usable (the reference would have exit(-1)'d on errors)
    assert int(QGTC.val2bit(torch.ones(8, 8, device="cuda"), 1, False, False)[0, 0].item()) == -16777216  # 0xFF000000


def test_non_default_stream_and_async(capi, oracle):
    M, K, N, b = 200, 700, 32, 2
    A, X, ba, bx, _ = make_operands(oracle, M, K, N, 1, b, seed=21)
    s = torch.cuda.Stream()
    da, dx = dev_words(ba), dev_words(bx)
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        outs = [capi.bitMM2Bit(da, dx, M, K, N, 1, b, b) for _ in range(8)]
    s.synchronize()
    want = oracle.bitMM2Bit(ba, bx, M, K, N, 1, b, b)
    for o in outs:
        assert np.array_equal(host_u32(o), want)


@pytest.mark.parametrize("H,W,E", [(759, 759, 11000), (130, 513, 900), (3, 3, 5), (2258, 2258, 40000), (64, 4100, 0)])
def test_edges2bit_equals_dense_pack(capi, oracle, H, W, E):
    """Edge list -> packed adjacency == val2bit of the dense 0/1 matrix (duplicates collapse, bad edges ignored)."""
    rng = np.random.default_rng(H + E)
    rows = rng.integers(0, H, E).astype(np.int32)
    cols = rng.integers(0, W, E).astype(np.int32)
    if E > 10:
        rows[:3], cols[:3] = rows[3:6], cols[3:6]           # duplicates
        rows[6], cols[7] = -1, W + 5                        # out of range: ignored
    A = np.zeros((H, W), np.float32)
    ok = (rows >= 0) & (rows < H) & (cols >= 0) & (cols < W)
    A[rows[ok], cols[ok]] = 1
    got = capi.edges2bit(torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), H, W)
    want = oracle.val2bit(A, 1, False, False)
    assert tuple(got.shape) == want.shape and np.array_equal(host_u32(got), want)

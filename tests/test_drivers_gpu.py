"""GPU test of the drivers: ClusterIter + the main_qgtc op sequences on a small synthetic graph, every
batch's output compared with the oracle chain on the same adjacency / features."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("run_gin", [False, True])
@pytest.mark.parametrize("bits", [1, 2])
def test_cluster_forward_matches_oracle(coracle, run_gin, bits):
    import qgtc_ppopp22_b200
    from qgtc_ppopp22_b200 import cluster as C
    from qgtc_ppopp22_b200 import main_qgtc as drv
    QGTC = qgtc_ppopp22_b200.load_extension()
    g = C.make_synthetic_graph("tiny", 60)
    par = C.shuffled_partitions(g, 60)
    it = C.ClusterIter("tiny", g, 60, 20, None, bit_width=bits, run_GIN=run_gin, ops=QGTC, par_li=par)
    assert len(it) == 3
    D, h, ncls = g.feat.shape[1], 16, 10
    W = drv.pack_weights(QGTC, D, h, ncls, bits, torch.device("cuda"))
    o = coracle
    oW = (o.val2bit(np.ones((D, h), np.float32), bits, True, False), o.val2bit(np.ones((h, h), np.float32), bits, True, False),
          o.val2bit(np.ones((h, ncls), np.float32), bits, True, True))
    fwd = drv.gin_forward if run_gin else drv.gcn_forward
    for cid, (cluster, param) in zip(it.batch_ids, it):
        n = param[0]
        got = fwd(QGTC, cluster.bit_A, cluster.bit_X, W, n, D, bits).cpu().numpy()
        rows, cols = it.batch_edges[cid]
        A = np.zeros((n, n), np.float32)
        A[rows.numpy(), cols.numpy()] = 1.0
        X = g.feat[torch.from_numpy(it.batch_nodes[cid])].numpy()
        bA, bX = o.val2bit(A, 1, False, False), o.val2bit(X, bits, True, False)
        assert np.array_equal(cluster.bit_A.cpu().numpy().view(np.uint32), bA)
        b = bits
        if run_gin:
            t = o.bitMM2Bit(bA, bX, n, n, D, 1, b, b)
            t = o.bitMM2Bit(t, oW[0], n, D, h, b, b, b)
            t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
            t = o.bitMM2Bit(t, oW[1], n, h, h, b, b, b)
            t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
            want = o.bitMM2Int(t, oW[2], n, h, ncls, b, b, False)
        else:
            t = o.bitMM2Bit(bX, oW[0], n, D, h, b, b, b)
            t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
            t = o.bitMM2Bit(t, oW[1], n, h, h, b, b, b)
            t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
            t = o.bitMM2Bit(t, oW[2], n, h, ncls, b, b, b)
            want = o.bitMM2Int(bA, t, n, n, h, 1, b, False)
        assert np.array_equal(got, want)


def test_epoch_runner_eager_and_graph():
    from qgtc_ppopp22_b200 import cluster as C
    from qgtc_ppopp22_b200 import main_qgtc as drv
    g = C.make_synthetic_graph("tiny", 60)
    a = drv.build_parser().parse_args(["--dataset", "tiny", "--psize", "60", "--batch-size", "20", "--n-epochs", "2", "--use_QGTC"])
    r1 = drv.run(a, graph=g, quiet=True)
    a.cuda_graph = True
    r2 = drv.run(a, graph=g, quiet=True)
    a.streams = 3
    r3 = drv.run(a, graph=g, quiet=True)
    assert r1["batches"] == 3 and r1["epoch_ms"] > 0 and r2["epoch_ms"] > 0 and r3["epoch_ms"] > 0
    assert r1["checksum"] == r2["checksum"] == r3["checksum"]
    # op-major (six grouped launches per epoch), eager and graph-captured, with and without the tile bitmaps
    a.streams, a.cuda_graph, a.grouped = 1, False, True
    r4 = drv.run(a, graph=g, quiet=True)
    a.cuda_graph, a.tile_bitmap = True, "on"
    r5 = drv.run(a, graph=g, quiet=True)
    assert r4["grouped"] and r5["grouped"] and r5["tile_bitmaps_used"]
    assert r4["checksum"] == r5["checksum"] == r1["checksum"]
    a.run_GIN, a.cuda_graph, a.grouped = True, False, False
    r6 = drv.run(a, graph=g, quiet=True)
    a.cuda_graph, a.grouped = True, True
    r7 = drv.run(a, graph=g, quiet=True)
    assert r6["checksum"] == r7["checksum"]
    # the logits gather (single rank: concatenation in batch order)
    a.gather_logits = True
    r8 = drv.run(a, graph=g, quiet=True)
    assert r8["logits_batches_gathered"] == 3 and r8["logits_checksum_all_batches"] != 0.0


@pytest.mark.parametrize("run_gin", [False, True])
def test_grouped_forward_equals_per_batch_forward(run_gin):
    """main_qgtc.grouped_forward (one launch per op over all batches) == gcn_forward / gin_forward batch by batch."""
    import qgtc_ppopp22_b200
    from qgtc_ppopp22_b200 import cluster as C
    from qgtc_ppopp22_b200 import main_qgtc as drv
    QGTC = qgtc_ppopp22_b200.load_extension()
    g = C.make_synthetic_graph("tiny", 60)
    par = C.shuffled_partitions(g, 60)
    for bits in (1, 2, 4):
        it = C.ClusterIter("tiny", g, 60, 20, None, bit_width=bits, run_GIN=run_gin, ops=QGTC, par_li=par, from_edges=True, tile_bitmaps=True)
        D = g.feat.shape[1]
        W = drv.pack_weights(QGTC, D, 16, 10, bits, torch.device("cuda"))
        fwd = drv.gin_forward if run_gin else drv.gcn_forward
        want = [fwd(QGTC, c.bit_A, c.bit_X, W, p[0], D, bits) for c, p in it]
        A, X, ns = [c.bit_A for c in it.cTensor_li], [c.bit_X for c in it.cTensor_li], [p[0] for p in it.cluster_param_li]
        for maps in ([], list(it.tile_maps)):
            for fused in (False, True):
                got = drv.grouped_forward(QGTC, run_gin, A, X, maps, W, ns, D, bits, fused=fused)
                assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want)), (bits, fused, bool(maps))
        assert 0.0 < it.nonzero_tile_fraction() <= 1.0


def test_cluster_iter_from_edges_matches_dense_path():
    import qgtc_ppopp22_b200
    from qgtc_ppopp22_b200 import cluster as C
    QGTC = qgtc_ppopp22_b200.load_extension()
    g = C.make_synthetic_graph("tiny", 60)
    par = C.shuffled_partitions(g, 60)
    a = C.ClusterIter("tiny", g, 60, 20, None, bit_width=2, ops=QGTC, par_li=par)
    b = C.ClusterIter("tiny", g, 60, 20, None, bit_width=2, ops=QGTC, par_li=par, from_edges=True)
    for (ca, _), (cb, _) in zip(a, b):
        assert torch.equal(ca.bit_A, cb.bit_A) and torch.equal(ca.bit_X, cb.bit_X)


def test_gcnconv_qnt_matches_integer_math(oracle):
    """The layer API with matching layouts: logits equal the exact integer pipeline on the quantized values."""
    from qgtc_ppopp22_b200.QGTC_conv import GCNConv_Qnt
    torch.manual_seed(0)
    n, D, h, C, wb, ab = 300, 40, 16, 10, 2, 3
    A = (torch.rand(n, n) < 0.03).float().cuda()
    X = (torch.rand(n, D) * 9 - 0.5).cuda()
    m = GCNConv_Qnt(D, h, C, w_bit=wb, act_bit=ab).cuda()
    with torch.no_grad():
        m.W_in.copy_(torch.rand(D, h) * 5 - 0.5)
        m.W_out.copy_(torch.rand(h, C) * 5 - 0.5)
    got = m(A, X).cpu().numpy()
    q = lambda t, b: (oracle.quantize_val(t.detach().cpu().numpy(), b) & ((1 << b) - 1)).astype(np.int64)
    An, Xq, W1, W2 = A.cpu().numpy().astype(np.int64), q(X, ab), q(m.W_in, wb), q(m.W_out, wb)
    mask = (1 << ab) - 1
    xw = oracle.requantize(Xq @ W1, ab) & mask
    hq = oracle.requantize(An @ xw, ab) & mask
    hw = oracle.requantize(hq @ W2, ab) & mask
    want = (An @ hw).astype(np.float32)
    assert got.shape == (n, C) and np.array_equal(got, want)


@pytest.mark.parametrize("bits", [4, 8])
def test_arxiv_shaped_chain_matches_oracle(coracle, bits):
    """One ogbn-arxiv-sized batch (n ~ 2258, D = 128, h = 16, C = 40: BASELINE configs[2]) through the GCN op sequence
    of main_qgtc.py:146-154 at 4 and 8 bits, every op defined on raw memory, against the C oracle."""
    import qgtc_ppopp22_b200
    from qgtc_ppopp22_b200 import main_qgtc as drv
    QGTC = qgtc_ppopp22_b200.load_extension()
    n, D, h, ncls, b = 2258, 128, 16, 40, bits
    rng = np.random.default_rng(100 + bits)
    A = (rng.random((n, n)) < 0.004).astype(np.float32)
    X = rng.standard_normal((n, D)).astype(np.float32) * 3
    dA, dX = torch.from_numpy(A).cuda(), torch.from_numpy(X).cuda()
    bit_A, bit_X = QGTC.val2bit(dA, 1, False, False), QGTC.val2bit(dX, b, True, False)
    W = drv.pack_weights(QGTC, D, h, ncls, b, torch.device("cuda"))
    got = drv.gcn_forward(QGTC, bit_A, bit_X, W, n, D, b).cpu().numpy()
    o = coracle
    oW = (o.val2bit(np.ones((D, h), np.float32), b, True, False), o.val2bit(np.ones((h, h), np.float32), b, True, False),
          o.val2bit(np.ones((h, ncls), np.float32), b, True, True))
    bA, bX = o.val2bit(A, 1, False, False), o.val2bit(X, b, True, False)
    assert np.array_equal(bit_A.cpu().numpy().view(np.uint32), bA) and np.array_equal(bit_X.cpu().numpy().view(np.uint32), bX)
    t = o.bitMM2Bit(bX, oW[0], n, D, h, b, b, b)
    t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
    t = o.bitMM2Bit(t, oW[1], n, h, h, b, b, b)
    t = o.bitMM2Bit(bA, t, n, n, h, 1, b, b)
    t = o.bitMM2Bit(t, oW[2], n, h, ncls, b, b, b)
    want = o.bitMM2Int(bA, t, n, n, h, 1, b, False)
    assert np.array_equal(got, want)


def test_extension_fused_op_falls_back_outside_the_fused_kernel(coracle):
    """QGTC.bitMM2Bit_grouped_fused with N = 100 (> 64: not the fused kernel) issues the two grouped launches itself: same results."""
    import qgtc_ppopp22_b200
    QGTC = qgtc_ppopp22_b200.load_extension()
    rng = np.random.default_rng(4)
    b, N, N2 = 2, 100, 24
    ns = [300, 129]
    W = torch.from_numpy(coracle.val2bit(rng.integers(0, 4, (N, N2)).astype(np.float32), b, True, False).view(np.int32)).cuda()
    A, X, want = [], [], []
    for n in ns:
        a = coracle.val2bit((rng.random((n, n)) < 0.05).astype(np.float32), 1, False, False)
        x = coracle.val2bit(rng.integers(0, 4, (n, N)).astype(np.float32), b, True, False)
        o1 = coracle.bitMM2Bit(a, x, n, n, N, 1, b, b)
        want.append(coracle.bitMM2Bit(o1, W.cpu().numpy().view(np.uint32), n, N, N2, b, b, b))
        A.append(torch.from_numpy(a.view(np.int32)).cuda())
        X.append(torch.from_numpy(x.view(np.int32)).cuda())
    for NN, (AA, XX, ww) in ((N, (A, X, want)),):
        got = QGTC.bitMM2Bit_grouped_fused(AA, XX, ns, ns, NN, 1, b, b, W, N2, b, b)
        for g, w in zip(got, ww):
            assert np.array_equal(g.cpu().numpy().view(np.uint32), w)

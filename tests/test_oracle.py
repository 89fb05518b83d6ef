"""CPU tests of the oracle itself: hand-derived known answers, numpy vs C
restatement, and the value-level identities the domain offers."""
import numpy as np
import pytest

from conftest import rand_adj, rand_features


def u32(a):
    return np.asarray(a).reshape(-1).view(np.uint32)


# ---------------------------------------------------------------- known answers

def test_kat_quantize(oracle):
    x = np.array([-0.3, 0.5, 1.5, 2.5, 3.6, 4.0, 4.1, np.nan, -0.0, np.inf, -np.inf, 3.49, 3.5], np.float32)
    #                1    0    2    2   4    4    3    0     0     3      1      3     4
    assert oracle.quantize_val(x, 2).tolist() == [1, 0, 2, 2, 4, 4, 3, 0, 0, 3, 1, 3, 4]
    # only bits 0..b-1 are packed: q == 2^b packs as 0
    w = oracle.val2bit(np.array([[3.6, 4.0, 3.0]], np.float32), 2, False, False)
    assert oracle.bit2val(w, 2, 1, 3, False).tolist() == [[0, 0, 3]]


def test_kat_row_layout(oracle):
    w = oracle.val2bit(np.ones((3, 3), np.float32), 2, False, False)
    assert w.shape == (16, 4) and w.dtype == np.uint32
    assert [int(v) for v in w[:3, 0]] == [0xE0000000] * 3
    assert int(w.astype(np.uint64).sum()) == 3 * 0xE0000000


def test_kat_col_layout(oracle):
    w = oracle.val2bit(np.ones((3, 3), np.float32), 2, True, False)
    assert w.shape == (8, 128)
    flat = w.reshape(-1)
    assert [int(flat[4 * c]) for c in range(3)] == [0xE0000000] * 3
    assert int((flat != 0).sum()) == 3


def test_kat_col8_out_of_bounds_rule(oracle):
    # C8 with P8(W) != P128(W) and bits >= 2: plane 1 lands beyond the allocation and is dropped
    x = np.full((16, 10), 3.0, np.float32)
    w = oracle.val2bit(x, 2, True, True)
    assert w.shape == (2 * 4, 16)
    flat = w.reshape(-1)
    assert (flat[: 10 * 4: 4] == 0xFFFF0000).all()      # plane 0, 10 columns, 16 rows each
    assert int((flat != 0).sum()) == 10                   # nothing of plane 1 survives
    # well-defined when P8(W) == P128(W)
    x = np.full((16, 128), 3.0, np.float32)
    assert (oracle.val2bit(x, 2, True, True) == oracle.val2bit(x, 2, True, False)).all()


@pytest.mark.parametrize("K,expect_bits", [(3, 3), (4, 0), (5, 3), (9, 3)])
def test_kat_all_ones_gemm(oracle, K, expect_bits):
    A, X = np.ones((3, K), np.float32), np.ones((K, 3), np.float32)
    ba, bx = oracle.val2bit(A, 1, False, False), oracle.val2bit(X, 2, True, False)
    assert (oracle.bitMM2Int(ba, bx, 3, K, 3, 1, 2, True) == K).all()
    out = oracle.bitMM2Bit(ba, bx, 3, K, 3, 1, 2, 2)
    assert out.shape == (2 * 8, 4)
    assert (oracle.bit2val(out, 2, 3, 3, False) == expect_bits).all()    # sat(K): K==2^ob packs as 0


def test_requantize_rule(oracle):
    acc = np.array([-5, -1, 0, 1, 3, 4, 5, 1 << 20], np.int64)
    assert oracle.requantize(acc, 2).tolist() == [1, 1, 0, 1, 3, 4, 3, 3]


def test_counters(oracle):
    A = np.zeros((16, 256), np.float32)
    A[0, 0] = 1          # tile (bx=0, i=0)
    A[9, 200] = 1        # tile (bx=1, i=1)
    ba = oracle.val2bit(A, 1, False, False)
    total, nz = oracle.tile_counts(ba, 16, 256, 24, 1, 2)
    assert total == 32 * 2 * 3 * 1 * 2 * 2
    assert nz == 32 * 3 * 2 * 2


# ---------------------------------------------------------------- numpy oracle == C oracle

CASES = [(3, 3, 3, 1, 2, 2), (37, 300, 21, 1, 3, 3), (64, 128, 16, 2, 2, 2), (5, 7, 3, 3, 8, 4),
         (130, 513, 70, 1, 2, 2), (17, 129, 9, 8, 8, 8), (1, 1, 1, 1, 1, 1)]


@pytest.mark.parametrize("M,K,N,b1,b2,ob", CASES)
def test_numpy_vs_c(oracle, coracle, M, K, N, b1, b2, ob):
    rng = np.random.default_rng(M * 1000 + K)
    A, X = rand_features(rng, (M, K), b1), rand_features(rng, (K, N), b2)
    for cm, ol in [(False, False), (True, False), (True, True)]:
        assert (oracle.val2bit(A, b1, cm, ol) == coracle.val2bit(A, b1, cm, ol)).all()
    ba, bx, bx8 = oracle.val2bit(A, b1, False, False), oracle.val2bit(X, b2, True, False), oracle.val2bit(X, b2, True, True)
    assert (oracle.bitMM2Bit(ba, bx, M, K, N, b1, b2, ob) == coracle.bitMM2Bit(ba, bx, M, K, N, b1, b2, ob)).all()
    assert (oracle.bitMM2Int(ba, bx, M, K, N, b1, b2, True) == coracle.bitMM2Int(ba, bx, M, K, N, b1, b2, True)).all()
    assert (oracle.bitMM2Int(ba, bx8, M, K, N, b1, b2, False) == coracle.bitMM2Int(ba, bx8, M, K, N, b1, b2, False)).all()
    for quirk in (False, True):
        assert (oracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob, quirk) == coracle.bitMM2Bit_col(ba, bx, M, K, N, b1, b2, ob, quirk)).all()
    assert oracle.tile_counts(ba, M, K, N, b1, b2) == coracle.tile_counts(ba, M, K, N, b1, b2)
    assert (oracle.bit2val(bx, b2, K, N, True) == coracle.bit2val(bx, b2, K, N, True)).all()
    # layout-mismatched operands are defined on raw memory with zero-extension
    assert (oracle.bitMM2Bit(bx, ba, K, M, N, b2, b1, ob) == coracle.bitMM2Bit(bx, ba, K, M, N, b2, b1, ob)).all()


# ---------------------------------------------------------------- domain properties

@pytest.mark.parametrize("M,K,N,b1,b2", [(37, 300, 21, 1, 3), (64, 256, 64, 4, 4), (9, 1000, 5, 8, 8)])
def test_value_level_identity(oracle, M, K, N, b1, b2):
    rng = np.random.default_rng(7)
    A, X = rand_features(rng, (M, K), b1, wild=False), rand_features(rng, (K, N), b2, wild=False)
    ba, bx = oracle.val2bit(A, b1, False, False), oracle.val2bit(X, b2, True, False)
    acc = oracle.bitgemm_acc(ba, bx, M, K, N, b1, b2, True)
    assert (acc == oracle.int_gemm_identity(A.astype(np.int64), X.astype(np.int64))).all()
    # pack / unpack round trip in every layout
    assert (oracle.bit2val(ba, b1, M, K, False) == A).all()
    assert (oracle.bit2val(bx, b2, K, N, True) == X).all()


def test_col_output_feeds_next_gemm(oracle):
    # intended bitMM2Bit_col: its output is a valid B operand (unitest.py:86-110 flow)
    rng = np.random.default_rng(3)
    n, D, h, b = 40, 50, 16, 2
    A, X, W = rand_adj(rng, n, n, 0.2), rand_features(rng, (n, D), b, False), rand_features(rng, (D, h), b, False)
    bA, bX, bW = oracle.val2bit(A, 1, False, False), oracle.val2bit(X, b, False, False), oracle.val2bit(W, b, True, False)
    xw_col = oracle.bitMM2Bit_col(bX, bW, n, D, h, b, b, b)
    xw = oracle.requantize(X.astype(np.int64) @ W.astype(np.int64), b) & ((1 << b) - 1)
    assert (oracle.bit2val(xw_col, b, n, h, True) == xw).all()
    out = oracle.bitMM2Int(bA, xw_col, n, n, h, 1, b, True)
    assert (out == (A.astype(np.int64) @ xw).astype(np.float32)).all()


def test_full_size_c_oracle_all_ones(coracle):
    # BASELINE.json config[1] headline shape, reference throughput inputs (2_7c_QGTC_GEMM_INT8.py:6-7)
    M = K = 4096
    N = 64
    ba = coracle.val2bit(np.ones((M, K), np.float32), 1, False, False)
    bx = coracle.val2bit(np.ones((K, N), np.float32), 2, True, False)
    assert (coracle.bitMM2Int(ba, bx, M, K, N, 1, 2, True) == K).all()
    out = coracle.bitMM2Bit(ba, bx, M, K, N, 1, 2, 2)
    assert (coracle.bit2val(out, 2, M, N, False) == 3).all()

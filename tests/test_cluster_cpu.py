"""CPU tests of the host-side driver logic: synthetic graphs, batch planning, sharding across ranks
(world_size 2 over gloo), and the torch restatement of the DGL baseline."""
import os
import random

import numpy as np
import pytest
import torch

from qgtc_ppopp22_b200 import cluster as C
from qgtc_ppopp22_b200 import dgl_baseline as B


def test_synthetic_graph_shapes_and_partitions():
    g = C.make_synthetic_graph("ppi", 1500)
    assert (g.num_nodes, g.num_edges, tuple(g.feat.shape), g.num_classes) == (56944, 818716, (56944, 50), 121)
    allnodes = np.concatenate(g.parts)
    assert allnodes.size == g.num_nodes and np.array_equal(np.sort(allnodes), np.arange(g.num_nodes))
    sizes = [len(p) for p in g.parts]
    assert max(sizes) - min(sizes) <= 1                                  # balanced, like METIS output
    # 1500 partitions, 20 per batch -> 75 batches of ~759 nodes
    counts = C.batch_node_counts(C.shuffled_partitions(g, 1500), 1500, 20)
    assert len(counts) == 75 and 750 <= min(counts) and max(counts) <= 770


def test_shuffle_is_rank_independent():
    g = C.make_synthetic_graph("tiny", 60)
    random.seed(123)
    a = C.shuffled_partitions(g, 60)
    random.seed(999)
    b = C.shuffled_partitions(g, 60)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_batches_covers_and_balances(world):
    rng = np.random.default_rng(0)
    costs = list(rng.integers(500, 3000, size=75).astype(float) ** 2)
    shards = C.shard_batches(costs, world)
    assert sorted(sum(shards, [])) == list(range(75))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) <= 1.25 * (sum(costs) / world)


def test_partition_cache_uses_the_reference_format(tmp_path):
    """sampler.py:56-65: ./datasets/{name}_{psize}.npy, an object array read with np.load(allow_pickle=True)."""
    g = C.make_synthetic_graph("tiny", 60)
    d = str(tmp_path) + "/"
    first = C.cached_partition_list("tiny", g, 60, cache_dir=d)
    fn = os.path.join(d, "tiny_60.npy")
    assert os.path.exists(fn)
    raw = np.load(fn, allow_pickle=True)                                 # exactly what the reference does
    assert raw.dtype == object and len(raw) == 60 and all(np.array_equal(a, b) for a, b in zip(raw, first))
    # a cache written by someone else (e.g. METIS partitions from the reference) is picked up unchanged
    other = np.empty(60, dtype=object)
    perm = np.random.default_rng(1).permutation(g.num_nodes)
    for i in range(60):
        other[i] = perm[i * 50:(i + 1) * 50].astype(np.int64)
    np.save(fn, other)
    again = C.cached_partition_list("tiny", g, 60, cache_dir=d)
    assert all(np.array_equal(a, b) for a, b in zip(again, other))
    assert len(C.cached_partition_list(None, g, 60, cache_dir=d)) == 60   # no name: no cache


def test_small_graph_batches():
    g = C.make_small_graph_batches(num_graphs=64)
    assert len(g.parts) == 64 and all(16 <= len(p) <= 64 for p in g.parts)
    part_of = np.empty(g.num_nodes, dtype=np.int64)
    for i, p in enumerate(g.parts):
        part_of[p] = i
    assert (part_of[g.src] == part_of[g.dst]).all()                       # block-diagonal


def test_baseline_models_match_dense_math():
    torch.manual_seed(0)
    n, d, h, c = 40, 7, 5, 3
    rows, cols = torch.randint(0, n, (200,)), torch.randint(0, n, (200,))
    A = B.batch_csr(rows, cols, n)
    dense = torch.zeros(n, n)
    dense.index_put_((cols, rows), torch.ones(200), accumulate=True)      # out[dst] += h[src]
    x = torch.randn(n, d)
    sage, gin = B.GraphSAGE(d, h, c), B.GIN(d, h, c)
    hh = x
    for lin in sage.layers:
        hh = torch.relu(lin(dense @ hh))
    assert torch.allclose(sage(A, x), hh, atol=1e-4)
    hh = x
    for lin in gin.layers:
        hh = lin(hh + dense @ hh)
    assert torch.allclose(gin(A, x), hh, atol=1e-4)


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = C.make_synthetic_graph("tiny", 60)
    par = C.shuffled_partitions(g, 60)
    counts = C.batch_node_counts(par, 60, 4)
    shards = C.shard_batches([float(n) * n for n in counts], world)
    mine = torch.zeros(len(counts), dtype=torch.int64)
    mine[shards[rank]] = 1
    dist.all_reduce(mine)                                                 # every batch owned by exactly one rank
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                              # the epoch time is the slowest rank's
    # the logits gather (the path's only data collective): every rank ends up with every batch's logits, in batch order
    logits = [torch.full((counts[b], 3), float(b)) + torch.arange(counts[b], dtype=torch.float32).view(-1, 1) for b in shards[rank]]
    full = C.gather_logits(logits, shards[rank], counts, world)
    ok = len(full) == len(counts) and all(
        tuple(full[b].shape) == (counts[b], 3) and float(full[b][0, 0]) == float(b) and float(full[b][-1, 2]) == float(b + counts[b] - 1)
        for b in range(len(counts)))
    q.put((rank, mine.tolist(), float(t.item()), [int(c) for c in counts], ok))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, owned, tmax, counts, gathered_ok in res:
        assert owned == [1] * len(counts)
        assert tmax == 11.0
        assert gathered_ok


class _OracleOps:
    """The two packing entry points ClusterIter needs, backed by the numpy oracle (CPU tensors)."""

    def __init__(self, oracle, batched):
        self.o, self.calls, self.batched_calls = oracle, 0, []
        if batched:
            self.val2bit_batched = self._batched

    def val2bit(self, x, nbits, col_major, output_layer):
        self.calls += 1
        return torch.from_numpy(self.o.val2bit(x.numpy(), nbits, col_major, output_layer).view(np.int32).copy())

    def _batched(self, xs, nbits, col_major, output_layer):
        self.batched_calls.append(len(xs))
        return [torch.from_numpy(self.o.val2bit(x.numpy(), nbits, col_major, output_layer).view(np.int32).copy()) for x in xs]


def test_cluster_iter_dense_route_batches_the_adjacency_packs(oracle):
    """Dense route (sampler.py:87-102): ops with val2bit_batched pack all batch adjacencies in one call (in batch order), ops without
    it (the reference's extension) one call per batch; the packed batches are identical either way and equal the oracle's pack of
    the induced dense adjacency."""
    g = C.make_synthetic_graph("tiny", 60)
    random.seed(2)
    par = C.shuffled_partitions(g, 60)
    plain, batched = _OracleOps(oracle, False), _OracleOps(oracle, True)
    a = C.ClusterIter("tiny", g, 60, 20, None, bit_width=2, device="cpu", ops=plain, par_li=par)
    b = C.ClusterIter("tiny", g, 60, 20, None, bit_width=2, device="cpu", ops=batched, par_li=par)
    assert len(a.cTensor_li) == len(b.cTensor_li) == 3
    assert batched.batched_calls == [3] and batched.calls == 3            # adjacencies: one batched call; features: one call per batch
    assert plain.calls == 6
    for i, (ta, tb) in enumerate(zip(a.cTensor_li, b.cTensor_li)):
        assert torch.equal(ta.bit_A, tb.bit_A) and torch.equal(ta.bit_X, tb.bit_X)
        nodes = a.batch_nodes[i]
        n = nodes.size
        loc = {int(v): k for k, v in enumerate(nodes)}
        A = np.zeros((n, n), np.float32)
        for s, d in zip(g.src, g.dst):
            if int(s) in loc and int(d) in loc:
                A[loc[int(s)], loc[int(d)]] = 1.0
        assert np.array_equal(ta.bit_A.numpy().view(np.uint32), oracle.val2bit(A, 1, False, False))

"""Batch construction for the Cluster-GCN / batched-GIN drivers (DGL-free).

Host-side mirror of the reference's ``sampler.py`` (ClusterIter, ClusterTensor) and
``partition_utils.py``: partitions -> batches of ``batch_size`` partitions -> induced subgraph ->
dense 0/1 adjacency -> ``QGTC.val2bit`` (sampler.py:76-106).  DGL, METIS and the datasets are not
available offline, so graphs are synthetic: a planted-partition graph whose partitions play the role
of METIS's output (balanced parts, mostly intra-part edges), with the reference's node counts, edge
counts, feature widths and class counts (BASELINE.json configs; dataset.py:64 uses randn features).

Only the packed tensors reach the GPU hot path; everything here is preparation outside the timed
region, exactly as in the reference (main_qgtc.py:79 builds the ClusterIter before start_time).
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np
import torch

# name -> (nodes, directed edges, feature dim, classes); BASELINE.json configs[0], [2] and SURVEY 8(d)
SHAPES = {
    "ppi": (56944, 818716, 50, 121),
    "ogbn-arxiv": (169343, 1166243, 128, 40),
    "Proteins": (43471, 162088, 29, 2),        # 0_7b_eval_QGTC_batched_GIN.py:13 row, TU-Proteins sized
    "tiny": (3000, 24000, 32, 7),
}


@dataclass
class SyntheticGraph:
    """Edge list + features + a partition of the nodes (what sampler.py gets from DGL + METIS)."""
    name: str
    num_nodes: int
    src: np.ndarray
    dst: np.ndarray
    feat: torch.Tensor          # [num_nodes, dim] float32, CPU
    num_classes: int
    parts: list                 # list of int64 arrays (node ids per partition), METIS stand-in

    @property
    def num_edges(self) -> int:
        return int(self.src.size)


def make_synthetic_graph(name: str = "ppi", psize: int = 1500, seed: int = 2, p_intra: float = 0.85,
                         num_nodes: int | None = None, num_edges: int | None = None, dim: int | None = None,
                         num_classes: int | None = None) -> SyntheticGraph:
    """Planted-partition graph with `psize` balanced parts; seeds follow main_qgtc.py:45-47
    (torch 3, numpy 2, random 2)."""
    n0, e0, d0, c0 = SHAPES.get(name, SHAPES["tiny"])
    n, e, d, c = num_nodes or n0, num_edges or e0, dim or d0, num_classes or c0
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n).astype(np.int64)
    bounds = np.linspace(0, n, psize + 1).astype(np.int64)
    parts = [np.sort(perm[bounds[i]:bounds[i + 1]]) for i in range(psize)]
    part_of = np.empty(n, dtype=np.int64)
    for i, p in enumerate(parts):
        part_of[p] = i
    # intra-part edges: both ends inside the source's part; the rest are uniform
    src = rng.integers(0, n, size=e, dtype=np.int64)
    intra = rng.random(e) < p_intra
    dst = rng.integers(0, n, size=e, dtype=np.int64)
    ps = part_of[src[intra]]
    lo, hi = bounds[ps], bounds[ps + 1]
    dst[intra] = perm[lo + (rng.random(int(intra.sum())) * (hi - lo)).astype(np.int64)]
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(n, d, generator=g)
    return SyntheticGraph(name, n, src, dst, feat, c, parts)


def make_small_graph_batches(num_graphs: int = 1024, graphs_per_batch: int = 32, dim: int = 29, num_classes: int = 2,
                             seed: int = 2) -> SyntheticGraph:
    """Batched-GIN input (BASELINE configs[3]): many small graphs (16-64 nodes, average degree ~4); a batch
    is a group of `graphs_per_batch` graphs, i.e. a block-diagonal adjacency.  Each small graph is one
    'partition', so the same ClusterIter builds the batches."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(16, 65, size=num_graphs)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(offs[-1])
    srcs, dsts, parts = [], [], []
    for i in range(num_graphs):
        k = int(sizes[i])
        m = 2 * k                                   # ~4 directed edges per node after symmetrisation
        a = rng.integers(0, k, size=m) + offs[i]
        b = rng.integers(0, k, size=m) + offs[i]
        srcs += [a, b]
        dsts += [b, a]
        parts.append(np.arange(offs[i], offs[i + 1], dtype=np.int64))
    g = torch.Generator().manual_seed(3)
    return SyntheticGraph("small-graphs", n, np.concatenate(srcs), np.concatenate(dsts), torch.randn(n, dim, generator=g),
                          num_classes, parts)


def get_partition_list(g: "SyntheticGraph", psize: int):
    """partition_utils.py:11-18: the list of node-id arrays, one per partition.  The reference asks METIS (dgl.transform.
    metis_partition); DGL is not installable offline, so the synthetic graph carries its planted partition, which plays METIS's
    output.  A real partitioner plugs in here."""
    return [np.asarray(p, dtype=np.int64) for p in list(g.parts)[:psize]]


def cached_partition_list(dn, g: "SyntheticGraph", psize: int, cache_dir: str = "./datasets/"):
    """sampler.py:56-65: partitions of a named dataset are cached as ``./datasets/{name}_{psize}.npy`` -- an object array of
    node-id arrays written with np.save and read back with np.load(allow_pickle=True), the reference's own format -- so a cache
    produced by the reference (METIS partitions of the real datasets) is picked up unchanged.  dn = None / '' disables the cache."""
    import os
    if not dn:
        return get_partition_list(g, psize)
    fn = os.path.join(cache_dir, dn + "_{}.npy".format(psize))
    if os.path.exists(fn):
        par = np.load(fn, allow_pickle=True)
        par_li = [np.asarray(p, dtype=np.int64) for p in par]
        n_nodes = sum(len(p) for p in par_li)
        if len(par_li) == psize and n_nodes <= g.num_nodes and (n_nodes == 0 or max(int(p.max()) for p in par_li if len(p)) < g.num_nodes):
            return par_li
        # a stale cache of another graph: fall through and rewrite it
    os.makedirs(cache_dir, exist_ok=True)
    par_li = get_partition_list(g, psize)
    arr = np.empty(len(par_li), dtype=object)
    for i, p_ in enumerate(par_li):
        arr[i] = p_
    np.save(fn, arr)
    return par_li


class ClusterTensor(torch.nn.Module):
    """sampler.py:12-19: a container of the two packed buffers that moves with .cuda()/.cpu()."""

    def __init__(self, bit_A: torch.Tensor, bit_X: torch.Tensor):
        super().__init__()
        self.register_buffer("bit_A", bit_A)
        self.register_buffer("bit_X", bit_X)

    def forward(self):
        pass


class ClusterIter:
    """sampler.py:21-149.  Same constructor arguments; ``g`` is a SyntheticGraph.

    For every batch: node set = `batch_size` consecutive partitions of the shuffled partition list
    (partition_utils.py:20-24), A = dense adjacency of the induced subgraph (sampler.py:80-89),
    bit_A = val2bit(A, 1, False, False), bit_X = val2bit(X, bit_width, True, False) (sampler.py:97-102).
    ``param`` = (A_size_0, A_size_1, X_size_0, X_size_1).

    Extensions (not in the reference): ``device`` keeps the packed batches resident on a GPU instead
    of bouncing them through host memory, and ``batch_ids`` restricts construction to this rank's
    shard of the batches.
    """

    def __init__(self, dn, g: SyntheticGraph, psize, batch_size, seed_nid=None, use_pp=False, regular=False,
                 bit_width=2, run_GIN=False, device=None, batch_ids=None, keep_on_cpu=False, ops=None, par_li=None,
                 from_edges=False, tile_bitmaps=False, partition_cache=False):
        self.g = g
        self.psize = psize
        self.batch_size = batch_size
        self.bit_width = bit_width
        self.regular = regular
        self.max = int(psize // batch_size)
        if par_li is None:
            self.par_li = cached_partition_list(dn if partition_cache else None, g, psize)    # sampler.py:56-65
            random.shuffle(self.par_li)                   # sampler.py:68 (random.seed(2) set by the driver)
        else:
            self.par_li = list(par_li)                    # already shuffled (multi-rank runs share one order)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.cTensor_li, self.cluster_param_li, self.batch_nodes = [], [], []
        self.tile_maps, self._tile_stats = [], [0, 0]      # per batch: qgtc_tile_bitmap of bit_A (zero-tile jumping); [non-zero, total] tiles
        self.batch_ids = list(range(self.max)) if batch_ids is None else list(batch_ids)
        if ops is None:
            import qgtc_ppopp22_b200
            ops = qgtc_ppopp22_b200.load_extension()
        self.ops = ops

        # group the edges by batch once (an edge belongs to a batch iff both ends are in it)
        batch_of = np.full(g.num_nodes, -1, dtype=np.int64)
        local_id = np.zeros(g.num_nodes, dtype=np.int64)
        for cid in range(self.max):
            nodes = np.concatenate(self.par_li[cid * batch_size:(cid + 1) * batch_size]).astype(np.int64)
            batch_of[nodes] = cid
            local_id[nodes] = np.arange(nodes.size)
            self.batch_nodes.append(nodes)
        eb = batch_of[g.src]
        keep = (eb >= 0) & (eb == batch_of[g.dst])
        order = np.argsort(eb[keep], kind="stable")
        es, ed, ebk = g.src[keep][order], g.dst[keep][order], eb[keep][order]
        starts = np.searchsorted(ebk, np.arange(self.max + 1))
        self.batch_edges = {}

        # Dense route (sampler.py:87-102: dense fp32 n x n adjacency -> val2bit): with this repo's ops the dense adjacencies are packed by
        # ONE launch per chunk of batches (val2bit_batched; chunks bounded to 1 GiB of fp32), with the reference's ops one call per batch.
        dense_bits = {}
        if not from_edges and hasattr(ops, "val2bit_batched"):
            pend, pend_bytes = [], 0

            def flush():
                nonlocal pend, pend_bytes
                if pend:
                    for (c, _), b in zip(pend, ops.val2bit_batched([a for _, a in pend], 1, False, False)):
                        dense_bits[c] = b
                pend, pend_bytes = [], 0

            for cid in self.batch_ids:
                n = int(self.batch_nodes[cid].size)
                if pend_bytes + 4 * n * n > (1 << 30):
                    flush()
                A = torch.zeros((n, n), dtype=torch.float32, device=self.device)
                r_ = torch.from_numpy(local_id[es[starts[cid]:starts[cid + 1]]]).to(self.device)
                c_ = torch.from_numpy(local_id[ed[starts[cid]:starts[cid + 1]]]).to(self.device)
                A[r_, c_] = 1.0                                         # duplicate edges collapse, as to_dense() of 0/1 after clip
                pend.append((cid, A))
                pend_bytes += 4 * n * n
            flush()

        for cid in self.batch_ids:
            nodes = self.batch_nodes[cid]
            n = int(nodes.size)
            rows = torch.from_numpy(local_id[es[starts[cid]:starts[cid + 1]]])
            cols = torch.from_numpy(local_id[ed[starts[cid]:starts[cid + 1]]])
            self.batch_edges[cid] = (rows, cols)
            X = g.feat[torch.from_numpy(nodes)].to(self.device)
            A = None
            if from_edges:
                # extension: pack the adjacency straight from the edge list (no dense fp32 n x n intermediate)
                from . import capi
                bit_A = capi.edges2bit(rows.to(self.device, torch.int32), cols.to(self.device, torch.int32), n, n)
            elif cid in dense_bits:
                bit_A = dense_bits.pop(cid)
            else:
                A = torch.zeros((n, n), dtype=torch.float32, device=self.device)
                A[rows.to(self.device), cols.to(self.device)] = 1.0     # duplicate edges collapse, as to_dense() of 0/1 after clip
                bit_A = ops.val2bit(A, 1, False, False)
            bit_X = ops.val2bit(X, self.bit_width, True, False)
            if tile_bitmaps:
                # built once, when the batch is packed (the reference re-tests every 8 x 128 tile inside every GEMM, kernel.h:574-591)
                from . import capi
                bm = capi.tile_bitmap(bit_A, n, n).view(-1)
                self.tile_maps.append(bm)
                self._tile_stats[0] += int(sum(bin(w & 0xFFFFFFFF).count("1") for w in bm.cpu().tolist()))
                self._tile_stats[1] += ((n + 7) // 8 * 8 + 127) // 128 * ((n + 127) // 128)
            if keep_on_cpu:
                bit_A, bit_X = bit_A.cpu(), bit_X.cpu()
            self.cTensor_li.append(ClusterTensor(bit_A, bit_X))
            self.cluster_param_li.append((n, n, n, int(X.size(1))))
            del A

    def nonzero_tile_fraction(self) -> float:
        """processed / total 128 x 128-bit adjacency tiles over this iterator's batches (what parse_counter.py:16 reports)."""
        return self._tile_stats[0] / self._tile_stats[1] if self._tile_stats[1] else 1.0

    def __len__(self):
        return len(self.cTensor_li)

    def __iter__(self):
        self.n = 0
        return self

    def __next__(self):
        if self.n < len(self.cTensor_li):
            item, param = self.cTensor_li[self.n], self.cluster_param_li[self.n]
            self.n += 1
            return item, param
        raise StopIteration


def shuffled_partitions(g: SyntheticGraph, psize: int, cache_name=None):
    """The partition order every rank must agree on (sampler.py:68 after main_qgtc.py:47's random.seed(2)).  `cache_name` turns
    on the reference's partition cache (./datasets/{name}_{psize}.npy, sampler.py:56-65)."""
    par_li = cached_partition_list(cache_name, g, psize)
    random.Random(2).shuffle(par_li)
    return par_li


def batch_node_counts(par_li, psize: int, batch_size: int):
    return [int(sum(len(p) for p in par_li[c * batch_size:(c + 1) * batch_size])) for c in range(int(psize // batch_size))]


def shard_batches(costs, world_size: int):
    """Static assignment of batches to ranks, balanced by estimated cost (longest-processing-time first).
    Returns a list of `world_size` lists of batch ids.  Batches are independent: no halo,
    no exchange inside the forward."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: loads[k])
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def gather_logits(per_batch, batch_ids, row_counts, world: int):
    """All ranks' per-batch logits -> the full list ordered by batch id, on every rank (the only data collective of
    the multi-GPU path, after the forward; NCCL over NVLink on GPUs, gloo in the CPU tests).
    per_batch: this rank's float32 [n_i, C] tensors, batch_ids: their global ids, row_counts: rows of EVERY batch (known to
    all ranks from the partition list).  Ranks pad their concatenated logits to the largest shard and all_gather once."""
    if world == 1:
        order = sorted(range(len(batch_ids)), key=lambda i: batch_ids[i])
        return [per_batch[i] for i in order]
    import torch.distributed as dist
    C_ = per_batch[0].shape[1] if per_batch else 0
    meta = torch.zeros(2, dtype=torch.int64)
    meta[0], meta[1] = sum(int(t.shape[0]) for t in per_batch), C_
    dev = per_batch[0].device if per_batch else torch.device("cpu")
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta.to(dev))
    rows_max = max(int(m[0]) for m in metas)
    C_ = max(int(m[1]) for m in metas)
    mine = torch.zeros((rows_max, C_), dtype=torch.float32, device=dev)
    if per_batch:
        mine[:int(meta[0])] = torch.cat(list(per_batch), dim=0)
    ids = torch.full((len(row_counts),), -1, dtype=torch.int64, device=dev)     # owner rank of every batch
    ids[torch.as_tensor(list(batch_ids), dtype=torch.int64, device=dev)] = dist.get_rank()
    owners = [torch.empty_like(ids) for _ in range(world)]
    dist.all_gather(owners, ids)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    out, cursor = [None] * len(row_counts), [0] * world
    for b in range(len(row_counts)):
        r = next(k for k in range(world) if int(owners[k][b]) == k)
        out[b] = bufs[r][cursor[r]:cursor[r] + row_counts[b]]
        cursor[r] += row_counts[b]
    return out

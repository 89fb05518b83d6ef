#!/usr/bin/env python3
"""Cluster-GCN / batched-GIN forward benchmark on the QGTC ops -- host-side mirror of the reference's
``main_qgtc.py`` (same command-line flags, same op sequences main_qgtc.py:128-154, same
"Avg. Epoch: %.3f ms" line), on DGL-free synthetic graphs of the reference's shapes.

Multi-GPU: batches are independent, so they are sharded across ranks (one process per GPU,
``torchrun --nproc-per-node N``); NCCL only reduces the timing (MAX over ranks) and, on request,
gathers a checksum of the logits.  ``--cuda-graph`` captures a rank's whole epoch once and replays it.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import qgtc_ppopp22_b200  # noqa: E402
from qgtc_ppopp22_b200 import cluster as C  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="ppi")
    p.add_argument("--gpu", type=int, default=0, help="gpu")
    p.add_argument("--n-epochs", type=int, default=20, help="number of training epochs")
    p.add_argument("--batch-size", type=int, default=20, help="batch size")
    p.add_argument("--psize", type=int, default=1500, help="number of partitions")
    p.add_argument("--dim", type=int, default=10, help="input dimension of each dataset (default=10)")
    p.add_argument("--n-hidden", type=int, default=16, help="number of hidden gcn units (default=16)")
    p.add_argument("--n-classes", type=int, default=10, help="number of classes (default=10)")
    p.add_argument("--n-layers", type=int, default=1, help="number of hidden gcn layers (default=1)")
    p.add_argument("--bit_width", type=int, default=2, help="bitwidth for QGTC quantization (default=2)")
    p.add_argument("--use-pp", action="store_true", help="whether to use precomputation")
    p.add_argument("--regular", action="store_true", help="whether to use DGL")
    p.add_argument("--run_GIN", action="store_true", help="whether to run GIN model")
    p.add_argument("--use_QGTC", action="store_true", help="whether to use QGTC")
    p.add_argument("--zerotile_jump", action="store_true", help="whether to profile zero-tile jumping")
    # extensions
    p.add_argument("--cuda-graph", action="store_true", help="capture one epoch in a CUDA graph and replay it")
    p.add_argument("--cpu-baseline", action="store_true", help="also time the torch restatement of the DGL model on the host")
    p.add_argument("--streams", type=int, default=1, help="run independent batches on this many CUDA streams (they share nothing)")
    p.add_argument("--grouped", action="store_true",
                   help="issue the epoch op-major: ONE grouped launch per op over all of this rank's batches (6 launches per epoch)")
    p.add_argument("--no-fuse", action="store_true", help="--grouped: one launch per op (six per epoch) instead of fusing each update GEMM into the aggregation before it")
    p.add_argument("--tile-bitmap", choices=["auto", "on", "off"], default="auto",
                   help="--grouped: pass the batches' tile bitmaps (zero-tile jumping); auto = when under half of the tiles are non-zero")
    p.add_argument("--min-region-ms", type=float, default=0.0, help="--cuda-graph: replay enough epochs for a timed region of at least this many ms")
    p.add_argument("--partition-cache", action="store_true", help="cache the partition list as ./datasets/{dataset}_{psize}.npy like sampler.py:56-65")
    p.add_argument("--gather-logits", action="store_true", help="after the timed epochs, all_gather every batch's logits to all ranks (NCCL)")
    p.add_argument("--dense-pack", action="store_true", help="build batches through the dense fp32 n x n adjacency as sampler.py:87-89 does "
                                                              "(default: packed straight from the edge list)")
    return p


def gcn_forward(QGTC, bit_A, bit_X, W, n, D, bw):
    """main_qgtc.py:146-154 (3-layer GCN; the layout mismatches of the reference chain are kept: every op is
    defined on raw memory)."""
    bit_W1, bit_W2, bit_W3, h1, h2, out = W
    o = QGTC.bitMM2Bit(bit_X, bit_W1, n, D, h1, bw, bw, bw)
    o = QGTC.bitMM2Bit(bit_A, o, n, n, h1, 1, bw, bw)
    o = QGTC.bitMM2Bit(o, bit_W2, n, h1, h2, bw, bw, bw)
    o = QGTC.bitMM2Bit(bit_A, o, n, n, h2, 1, bw, bw)
    o = QGTC.bitMM2Bit(o, bit_W3, n, h2, out, bw, bw, bw)
    return QGTC.bitMM2Int(bit_A, o, n, n, h2, 1, bw, False)


def gin_forward(QGTC, bit_A, bit_X, W, n, D, bw):
    """main_qgtc.py:131-138."""
    bit_W1, bit_W2, bit_W3, h1, h2, out = W
    o = QGTC.bitMM2Bit(bit_A, bit_X, n, n, D, 1, bw, bw)
    o = QGTC.bitMM2Bit(o, bit_W1, n, D, h1, bw, bw, bw)
    o = QGTC.bitMM2Bit(bit_A, o, n, n, h1, 1, bw, bw)
    o = QGTC.bitMM2Bit(o, bit_W2, n, h1, h2, bw, bw, bw)
    o = QGTC.bitMM2Bit(bit_A, o, n, n, h2, 1, bw, bw)
    return QGTC.bitMM2Int(o, bit_W3, n, h2, out, bw, bw, False)


def grouped_forward(QGTC, run_gin, A, X, tile_maps, W, ns, D, bw, fused=True):
    """The same six ops for ALL batches, op-major: op i of every batch is one grouped launch (QGTC.bitMM2*_grouped), so an
    epoch is six kernels instead of six per batch.  Batches are independent and every op is the reference's op on the same
    raw memory, so each batch's result is bit-identical to gcn_forward / gin_forward (main_qgtc.py:128-154)."""
    bit_W1, bit_W2, bit_W3, h1, h2, out = W
    B = len(A)
    W1, W2, W3 = [bit_W1] * B, [bit_W2] * B, [bit_W3] * B
    # static_operands: the adjacency (X1 of the aggregations, with its tile bitmap) and the weights (X2 of the updates) are not
    # produced by the previous op, so the kernel may load them while that op is still draining (programmatic dependent launch)
    SA, SW = 1, 2
    if fused:
        # every update GEMM (o.W: K = hidden <= 64) rides in the epilogue of the aggregation before it (QGTC.bitMM2Bit_grouped_fused):
        # the requantized tile goes straight back into tensor memory as the second GEMM's A operand.  GCN: 4 launches, GIN: 3.
        if run_gin:
            o = QGTC.bitMM2Bit_grouped_fused(A, X, ns, ns, D, 1, bw, bw, bit_W1, h1, bw, bw, False, False, tile_maps, SA | SW)
            o = QGTC.bitMM2Bit_grouped_fused(A, o, ns, ns, h1, 1, bw, bw, bit_W2, h2, bw, bw, False, False, tile_maps, SA)
            return QGTC.bitMM2Bit_grouped_fused(A, o, ns, ns, h2, 1, bw, bw, bit_W3, out, bw, bw, True, False, tile_maps, SA)
        o = QGTC.bitMM2Bit_grouped(X, W1, ns, [D] * B, h1, bw, bw, bw, [], SA | SW)
        o = QGTC.bitMM2Bit_grouped_fused(A, o, ns, ns, h1, 1, bw, bw, bit_W2, h2, bw, bw, False, False, tile_maps, SA)
        o = QGTC.bitMM2Bit_grouped_fused(A, o, ns, ns, h2, 1, bw, bw, bit_W3, out, bw, bw, False, False, tile_maps, SA)
        return QGTC.bitMM2Int_grouped(A, o, ns, ns, h2, 1, bw, False, tile_maps, SA)
    if run_gin:
        o = QGTC.bitMM2Bit_grouped(A, X, ns, ns, D, 1, bw, bw, tile_maps, SA | SW)
        o = QGTC.bitMM2Bit_grouped(o, W1, ns, [D] * B, h1, bw, bw, bw, [], SW)
        o = QGTC.bitMM2Bit_grouped(A, o, ns, ns, h1, 1, bw, bw, tile_maps, SA)
        o = QGTC.bitMM2Bit_grouped(o, W2, ns, [h1] * B, h2, bw, bw, bw, [], SW)
        o = QGTC.bitMM2Bit_grouped(A, o, ns, ns, h2, 1, bw, bw, tile_maps, SA)
        return QGTC.bitMM2Int_grouped(o, W3, ns, [h2] * B, out, bw, bw, False, [], SW)
    o = QGTC.bitMM2Bit_grouped(X, W1, ns, [D] * B, h1, bw, bw, bw, [], SA | SW)
    o = QGTC.bitMM2Bit_grouped(A, o, ns, ns, h1, 1, bw, bw, tile_maps, SA)
    o = QGTC.bitMM2Bit_grouped(o, W2, ns, [h1] * B, h2, bw, bw, bw, [], SW)
    o = QGTC.bitMM2Bit_grouped(A, o, ns, ns, h2, 1, bw, bw, tile_maps, SA)
    o = QGTC.bitMM2Bit_grouped(o, W3, ns, [h2] * B, out, bw, bw, bw, [], SW)
    return QGTC.bitMM2Int_grouped(A, o, ns, ns, h2, 1, bw, False, tile_maps, SA)


def pack_weights(QGTC, feat_size, hidden, output, bw, device):
    """main_qgtc.py:100-110: all-ones weights, W3 in the output-layer (C8) format."""
    W_1 = torch.ones((feat_size, hidden), device=device)
    W_2 = torch.ones((hidden, hidden), device=device)
    W_3 = torch.ones((hidden, output), device=device)
    return (QGTC.val2bit(W_1, bw, True, False), QGTC.val2bit(W_2, bw, True, False), QGTC.val2bit(W_3, bw, True, True),
            hidden, hidden, output)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def run(args, graph=None, quiet=False, ops=None):
    """`ops` replaces the op module (default: this package's QGTC extension) -- e.g. the reference's own extension rebuilt for
    sm_100a, for a same-box reference epoch (bench.py --impl reference)."""
    torch.manual_seed(3)
    np.random.seed(2)
    random.seed(2)
    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(args.gpu)
    device = torch.device("cuda", torch.cuda.current_device())
    QGTC = ops if ops is not None else qgtc_ppopp22_b200.load_extension()
    ours = getattr(QGTC, "__backend__", None) == "qgtc_b200"

    if graph is None:
        if args.run_GIN and args.dataset in ("small-graphs", "Proteins-batched"):
            graph = C.make_small_graph_batches()
            args.psize, args.batch_size = len(graph.parts), 32
        else:
            known = args.dataset in C.SHAPES
            graph = C.make_synthetic_graph(args.dataset, args.psize, dim=None if known else args.dim)
    par_li = C.shuffled_partitions(graph, args.psize, args.dataset if getattr(args, "partition_cache", False) and (rank == 0 or world == 1) else None)
    if world > 1 and getattr(args, "partition_cache", False):      # rank 0 wrote / read the cache; the others read it after the barrier
        import torch.distributed as dist
        dist.barrier()
        par_li = C.shuffled_partitions(graph, args.psize, args.dataset)
    counts = C.batch_node_counts(par_li, args.psize, args.batch_size)
    shards = C.shard_batches([float(n) * n for n in counts], world)      # dominant cost ~ n^2 * hidden
    it = C.ClusterIter(args.dataset, graph, args.psize, args.batch_size, None, use_pp=False, regular=args.regular,
                       bit_width=args.bit_width, run_GIN=args.run_GIN, device=device, batch_ids=shards[rank], ops=QGTC, par_li=par_li,
                       from_edges=ours and not getattr(args, "dense_pack", False), tile_bitmaps=ours)

    feat_size = graph.feat.shape[1]
    bw = args.bit_width
    torch.cuda.synchronize()
    start_time = time.time()                                              # main_qgtc.py:96
    W = pack_weights(QGTC, feat_size, args.n_hidden, args.n_classes, bw, device)
    fwd = gin_forward if args.run_GIN else gcn_forward
    last = None

    side = [torch.cuda.Stream() for _ in range(max(0, args.streams - 1))]

    def one_epoch():
        """All batches of this rank once.  Batches are independent, so with --streams K they are issued
        round-robin on K streams (forked from / joined to the current stream, which also makes the
        pattern capturable as parallel branches of one CUDA graph)."""
        out = None
        if not side:
            for cluster, param in it:
                out = fwd(QGTC, cluster.bit_A, cluster.bit_X, W, param[0], param[3], bw)
            return out
        cur = torch.cuda.current_stream()
        streams = [cur] + side
        for st in side:
            st.wait_stream(cur)
        for i, (cluster, param) in enumerate(it):
            with torch.cuda.stream(streams[i % len(streams)]):
                out = fwd(QGTC, cluster.bit_A, cluster.bit_X, W, param[0], param[3], bw)
        for st in side:
            cur.wait_stream(st)
        return out

    grouped = bool(getattr(args, "grouped", False)) and ours and len(it) > 0
    if grouped:
        gA = [c.bit_A for c in it.cTensor_li]
        gX = [c.bit_X for c in it.cTensor_li]
        gns = [p[0] for p in it.cluster_param_li]
        mode = getattr(args, "tile_bitmap", "auto")
        use_maps = it.tile_maps and (mode == "on" or (mode == "auto" and it.nonzero_tile_fraction() < 0.5))
        gmaps = list(it.tile_maps) if use_maps else []

        def one_epoch():  # noqa: F811  (op-major: six grouped launches)
            res = grouped_forward(QGTC, args.run_GIN, gA, gX, gmaps, W, gns, feat_size, bw, fused=not getattr(args, "no_fuse", False))
            return res[-1]

    if args.zerotile_jump:
        for cluster, param in it:
            QGTC.bitMM2Bit_base_cnt(cluster.bit_X, W[0], param[2], param[3], W[3], bw, bw, bw)
            QGTC.bitMM2Bit_zerojump_cnt(cluster.bit_X, W[0], param[2], param[3], W[3], bw, bw, bw)
        return {"epoch_ms": None}

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_timed = args.n_epochs
    if args.cuda_graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            one_epoch()                                                   # warm-up outside capture
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph_obj = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_obj):
            last = one_epoch()
        for _ in range(3):
            graph_obj.replay()
        torch.cuda.synchronize()
        min_region_ms = float(getattr(args, "min_region_ms", 0.0) or 0.0)
        if min_region_ms > 0:                                             # size the timed region: replays are tens of microseconds
            e0.record()
            for _ in range(20):
                graph_obj.replay()
            e1.record()
            torch.cuda.synchronize()
            est = max(e0.elapsed_time(e1) / 20, 1e-3)
            n_timed = max(n_timed, int(min_region_ms / est) + 1)
            if world > 1:                                                 # every rank replays the same number of epochs
                import torch.distributed as dist
                t = torch.tensor([n_timed], dtype=torch.int64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                n_timed = int(t.item())
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        start_time = time.time()
        e0.record()
        for _ in range(n_timed):
            graph_obj.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(n_timed):
            last = one_epoch()
        e1.record()
    torch.cuda.synchronize()
    ms = (time.time() - start_time) * 1000 / n_timed                      # the reference's convention (main_qgtc.py:157-159)
    ms_dev = e0.elapsed_time(e1) / n_timed                                # device time of the same epochs (CUDA events on the launching stream)
    rank_ms = [ms_dev]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        allr = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(allr, torch.tensor([ms_dev], dtype=torch.float64, device=device))
        rank_ms = [float(x.item()) for x in allr]
        ms_dev = max(rank_ms)
    result = {"epoch_ms": ms, "epoch_ms_device": ms_dev, "epochs_timed": n_timed, "rank_epoch_ms_device": [round(x, 5) for x in rank_ms],
              "batches": len(it), "world": world, "grouped": grouped,
              "tile_bitmaps_used": bool(grouped and gmaps), "nonzero_tile_fraction": it.nonzero_tile_fraction() if it.tile_maps else None,
              "checksum": float(last.sum().item()) if last is not None else 0.0, "nodes_per_batch": int(np.mean(counts))}
    if getattr(args, "gather_logits", False):
        # one more (untimed) forward that keeps every batch's logits, then the only data collective of the path
        if grouped:
            mine = grouped_forward(QGTC, args.run_GIN, gA, gX, gmaps, W, gns, feat_size, bw, fused=not getattr(args, "no_fuse", False))
        else:
            mine = [fwd(QGTC, c.bit_A, c.bit_X, W, p_[0], p_[3], bw) for c, p_ in it]
        full = C.gather_logits(list(mine), list(it.batch_ids), counts, world)
        result["logits_checksum_all_batches"] = float(sum(float(t.double().sum().item()) for t in full))
        result["logits_batches_gathered"] = len(full)
    if args.cpu_baseline and rank == 0:
        from qgtc_ppopp22_b200 import dgl_baseline
        full = it if world == 1 else C.ClusterIter(args.dataset, graph, args.psize, args.batch_size, None, bit_width=bw,
                                                   run_GIN=args.run_GIN, device=device, ops=QGTC, par_li=par_li)
        cpu_ms, cores = dgl_baseline.cpu_epoch_ms(graph, full.batch_edges, full.batch_nodes, args.n_hidden, graph.num_classes,
                                                  run_gin=args.run_GIN)
        result.update({"cpu_epoch_ms": cpu_ms, "cpu_cores": cores})
    if rank == 0 and not quiet:
        print("Avg. Epoch: {:.3f} ms".format(ms))                         # main_qgtc.py:159, parsed by parse_time.py
        if "cpu_epoch_ms" in result:
            print("CPU (torch restatement of the DGL model, %d cores) Avg. Epoch: %.3f ms" % (result["cpu_cores"], result["cpu_epoch_ms"]))
    return result


if __name__ == "__main__":
    a = build_parser().parse_args()
    print(a)
    run(a)

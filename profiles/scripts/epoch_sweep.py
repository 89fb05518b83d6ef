#!/usr/bin/env python3
"""BASELINE configs[2] and [3]: Cluster-GCN on a synthetic ogbn-arxiv-shaped graph at 1/2/4/8 bits and batched
GIN (2-bit) over 1024 small graphs; epoch ms eager / CUDA graph / CUDA graph on 8 streams.  Run under torchrun
for N > 1 (batches are sharded across ranks)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import cluster as C  # noqa: E402
from qgtc_ppopp22_b200 import main_qgtc as drv  # noqa: E402

if __name__ == "__main__":
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    runs = []
    g = C.make_synthetic_graph("ogbn-arxiv", 1500)
    for bits in (1, 2, 4, 8):
        runs.append(("cluster_gcn ogbn-arxiv-shaped", g, ["--dataset", "ogbn-arxiv", "--n-classes", "40", "--bit_width", str(bits)]))
    gp = C.make_synthetic_graph("ppi", 1500)
    runs.append(("cluster_gcn ppi-shaped", gp, ["--dataset", "ppi", "--n-classes", "121", "--bit_width", "2"]))
    sg = C.make_small_graph_batches()
    runs.append(("batched_gin 1024 small graphs", sg, ["--dataset", "small-graphs", "--run_GIN", "--n-hidden", "64", "--n-classes", "2",
                                                        "--psize", "1024", "--batch-size", "32", "--bit_width", "2"]))
    for name, graph, extra in runs:
        res = {"config": name, "n_gpus": world, "args": " ".join(extra)}
        for label, more in (("eager", []), ("cuda_graph", ["--cuda-graph"]), ("cuda_graph_8_streams", ["--cuda-graph", "--streams", "8"])):
            a = drv.build_parser().parse_args(extra + ["--use_QGTC", "--n-epochs", "10"] + more)
            r = drv.run(a, graph=graph, quiet=True)
            res["epoch_ms_" + label] = round(r["epoch_ms"], 3)
            res["batches_on_rank0"], res["nodes_per_batch"] = r["batches"], r["nodes_per_batch"]
        if rank == 0:
            print(json.dumps(res), flush=True)

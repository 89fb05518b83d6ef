#!/usr/bin/env python3
"""Epoch ms of the drivers per launch mode: per-batch eager / one CUDA graph / graph + 16 streams (round 1) and the op-major
grouped mode (six launches per epoch, eager and graph-captured).  CUDA-event timing of >= 100 ms regions is in bench.py; this
sweep uses the driver's own wall clock over many epochs.  Usage: epoch_modes.py [ppi|ogbn-arxiv|gin] [bits ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import cluster as C  # noqa: E402
from qgtc_ppopp22_b200 import main_qgtc as drv  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ppi"
bits = [int(b) for b in sys.argv[2:]] or [2]
if which == "gin":
    g = C.make_small_graph_batches()
    base = ["--dataset", "small-graphs", "--run_GIN", "--n-hidden", "64", "--n-classes", "2", "--use_QGTC", "--psize", str(len(g.parts)), "--batch-size", "32"]
else:
    g = C.make_synthetic_graph(which, 1500)
    base = ["--dataset", which, "--use_QGTC", "--n-classes", str(g.num_classes)]
MODES = {"eager": [], "graph": ["--cuda-graph"], "graph_16streams": ["--cuda-graph", "--streams", "16"],
         "grouped_eager": ["--grouped"], "grouped_graph": ["--grouped", "--cuda-graph"], "grouped_graph_unfused": ["--grouped", "--cuda-graph", "--no-fuse"],
         "grouped_graph_nomaps": ["--grouped", "--cuda-graph", "--tile-bitmap", "off"], "grouped_graph_maps": ["--grouped", "--cuda-graph", "--tile-bitmap", "on"]}
for b in bits:
    row = {"graph": which, "bits": b}
    for name, extra in MODES.items():
        a = drv.build_parser().parse_args(base + ["--bit_width", str(b), "--n-epochs", "200" if "graph" in name else "20"] + extra)
        r = drv.run(a, graph=g, quiet=True)
        row[name + "_ms"] = round(r["epoch_ms"], 4)
        row["checksum_" + name] = r["checksum"]
        row["nonzero_tile_fraction"] = r["nonzero_tile_fraction"]
    cs = {v for k, v in row.items() if k.startswith("checksum_")}
    row = {k: v for k, v in row.items() if not k.startswith("checksum_")}
    row["all_modes_same_checksum"] = len(cs) == 1
    print(json.dumps(row), flush=True)

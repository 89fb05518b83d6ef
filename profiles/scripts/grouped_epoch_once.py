#!/usr/bin/env python3
"""A few op-major (grouped) epochs of a driver config, eager launches -- the target of ncu launch lists / captures.
Usage: grouped_epoch_once.py [ppi|ogbn-arxiv|gin] [bits] [epochs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qgtc_ppopp22_b200 import cluster as C  # noqa: E402
from qgtc_ppopp22_b200 import main_qgtc as drv  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ppi"
bits = sys.argv[2] if len(sys.argv) > 2 else "2"
epochs = sys.argv[3] if len(sys.argv) > 3 else "3"
if which == "gin":
    g = C.make_small_graph_batches()
    base = ["--dataset", "small-graphs", "--run_GIN", "--n-hidden", "64", "--n-classes", "2", "--use_QGTC", "--psize", str(len(g.parts)), "--batch-size", "32"]
else:
    g = C.make_synthetic_graph(which, 1500)
    base = ["--dataset", which, "--use_QGTC", "--n-classes", str(g.num_classes)]
a = drv.build_parser().parse_args(base + ["--bit_width", bits, "--n-epochs", epochs, "--grouped"] + sys.argv[4:])
print(drv.run(a, graph=g, quiet=True))

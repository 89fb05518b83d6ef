"""The reference's DGL baseline models restated in plain PyTorch (DGL is not installed offline).

modules.py:9-45  GraphSAGE: 3 x ( sum-aggregate over in-edges -> nn.Linear -> ReLU ), ReLU after the last too.
modules.py:65-99 GIN:       3 x  Linear( (1 + eps) * h + sum_{nbr} h ), eps = 0 (GINConv 'sum', learn_eps=False).
Used as the reported CPU baseline of the Cluster-GCN / batched-GIN epoch (cluster_gcn_dgl.py:97-105:
autograd left on, no backward).  Aggregation is a CSR sparse-dense product on the batch's adjacency.
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class GraphSAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers=1):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(in_feats, n_hidden), nn.Linear(n_hidden, n_hidden), nn.Linear(n_hidden, n_classes)])

    def forward(self, A_csr, h):
        for lin in self.layers:
            h = F.relu(lin(torch.sparse.mm(A_csr, h)))
        return h


class GIN(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers=3):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(input_dim, hidden_dim), nn.Linear(hidden_dim, hidden_dim), nn.Linear(hidden_dim, output_dim)])

    def forward(self, A_csr, h):
        for lin in self.layers:
            h = lin(h + torch.sparse.mm(A_csr, h))
        return h


def batch_csr(rows: torch.Tensor, cols: torch.Tensor, n: int) -> torch.Tensor:
    """in-edge aggregation matrix: out[dst] += h[src]  (update_all(copy_src, sum))"""
    import warnings
    idx = torch.stack([cols, rows])            # row = destination, column = source
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = torch.sparse_coo_tensor(idx, torch.ones(rows.numel()), (n, n)).coalesce()
        return a.to_sparse_csr()


def cpu_epoch_ms(graph, batch_edges: dict, batch_nodes: list, n_hidden: int, n_classes: int, run_gin: bool = False,
                 epochs: int = 3, warmup: int = 1) -> tuple[float, int]:
    """Average epoch time (ms) of the baseline forward over the given batches on the host cores."""
    import os
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; the baseline gets every core
    torch.manual_seed(3)
    in_feats = graph.feat.shape[1]
    model = GIN(in_feats, n_hidden, n_classes) if run_gin else GraphSAGE(in_feats, n_hidden, n_classes)
    items = []
    for cid, (rows, cols) in batch_edges.items():
        nodes = torch.from_numpy(batch_nodes[cid])
        items.append((batch_csr(rows, cols, nodes.numel()), graph.feat[nodes]))
    for _ in range(warmup):
        for a, x in items:
            model(a, x)
    t0 = time.perf_counter()
    for _ in range(epochs):
        for a, x in items:
            model(a, x)
    return (time.perf_counter() - t0) * 1e3 / epochs, torch.get_num_threads()

"""Working version of the reference's layer API (``QGTC_conv.py``: Aggregation_Qnt, GCNConv_Qnt, GCNConv).

The reference file cannot run (4-argument ``bitMM2Bit`` calls, ``super(GCNConv, self)`` in the wrong class,
QGTC_conv.py:15-21,40) and its drivers chain ops with mismatched layouts, so their logits are not meaningful
(SURVEY.md section 0).  This module keeps the same class names and the same two-GEMM structure per layer
(update  X.W  then aggregate  A.(XW)), but with layouts that match, following the only flow the reference
exercises correctly (unitest.py:86-110):

    X  --val2bit(row)-->  R(b, n, D)             A operand of the update GEMM
    W  --val2bit(col)-->  C128(b, D, h)          B operand
    XW = bitMM2Bit_col(X, W)  -> C128(b, n, h)   requantized, column-major: a valid B operand
    hidden:  bitMM2Bit(A, XW) -> R(b, n, h)      row-major: the next layer's A operand
    output:  bitMM2Int(A, XW, pad_128=True) -> float [n, C]

Every product is the exact integer GEMM of the quantized values (value-level identity), so the module can be
checked against int64 matmuls of the quantized operands -- tests/test_drivers_gpu.py does.
No backward, as in the reference (QGTC_conv.py:24-27).
"""
from __future__ import annotations

import torch

import qgtc_ppopp22_b200


def _ops():
    return qgtc_ppopp22_b200.load_extension()


class Aggregation_Qnt(torch.autograd.Function):
    """out = A . requant(X . W) on packed operands (QGTC_conv.py:9-22).  `bit_X` is R(act_bit, n, D), `bit_W` is
    C128(w_bit, D, h), `bit_A` is R(1, n, n).  hidden/input -> packed R(act_bit, n, h); output -> float [n, h]."""

    @staticmethod
    def forward(ctx, bit_A, bit_X, bit_W, n, D, h, act_bit, w_bit, input=False, hidden=False, output=False):
        assert input + hidden + output == 1
        QGTC = _ops()
        bit_xw = QGTC.bitMM2Bit_col(bit_X, bit_W, n, D, h, act_bit, w_bit, act_bit)        # C128(act_bit, n, h)
        if input or hidden:
            return QGTC.bitMM2Bit(bit_A, bit_xw, n, n, h, 1, act_bit, act_bit)             # R(act_bit, n, h)
        return QGTC.bitMM2Int(bit_A, bit_xw, n, n, h, 1, act_bit, True)                    # float [n, h]

    @staticmethod
    def backward(ctx, d_output):
        return (None,) * 11


class GCNConv_Qnt(torch.nn.Module):
    """Two-layer quantized GCN (QGTC_conv.py:38-78): weights packed once, adjacency and features per call."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers=2, w_bit=2, act_bit=3):
        super().__init__()
        self.W_in = torch.nn.Parameter(torch.randn(input_dim, hidden_dim))
        self.W_out = torch.nn.Parameter(torch.randn(hidden_dim, output_dim))
        self.dims = (input_dim, hidden_dim, output_dim)
        self.w_bit, self.act_bit = w_bit, act_bit
        self.bit_W_in = None
        self.bit_W_out = None

    def weight_Qnt(self):
        QGTC = _ops()
        self.bit_W_in = QGTC.val2bit(self.W_in.detach().contiguous(), self.w_bit, True, False)
        self.bit_W_out = QGTC.val2bit(self.W_out.detach().contiguous(), self.w_bit, True, False)

    def A_Qnt(self, A):
        return _ops().val2bit(A.contiguous(), 1, False, False)

    def X_Qnt(self, X):
        return _ops().val2bit(X.contiguous(), self.act_bit, False, False)

    def forward(self, A, X):
        """A: dense [n, n] 0/1 adjacency of a subgraph, X: [n, input_dim] embedding; returns float [n, output_dim]."""
        if self.bit_W_in is None:
            self.weight_Qnt()
        n = A.size(0)
        D, h, C = self.dims
        bit_A, bit_X = self.A_Qnt(A), self.X_Qnt(X)
        bit_h = Aggregation_Qnt.apply(bit_A, bit_X, self.bit_W_in, n, D, h, self.act_bit, self.w_bit, True, False, False)
        return Aggregation_Qnt.apply(bit_A, bit_h, self.bit_W_out, n, h, C, self.act_bit, self.w_bit, False, False, True)


class Aggregation(torch.autograd.Function):
    """fp32 counterpart (QGTC_conv.py:82-91)."""

    @staticmethod
    def forward(ctx, A, X, W):
        return torch.mm(A, torch.mm(X, W))

    @staticmethod
    def backward(ctx, d_output):
        return None, None, None


class GCNConv(torch.nn.Module):
    """QGTC_conv.py:105-121."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers=2):
        super().__init__()
        self.W_in = torch.nn.Parameter(torch.randn(input_dim, hidden_dim))
        self.W_out = torch.nn.Parameter(torch.randn(hidden_dim, output_dim))

    def forward(self, A, X):
        return Aggregation.apply(A, Aggregation.apply(A, X, self.W_in), self.W_out)

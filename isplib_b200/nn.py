"""Minimal GNN layers that call ``torch_sparse.matmul(adj_t, x, reduce)`` exactly the way
PyG's GCNConv / SAGEConv / GINConv do when given a SparseTensor (PyG is not installed in
this image).  They are the CALLERS of the hot path in the reference's benchmark scripts:

* GCN   /root/reference/tests/cpu/gcn-sparse.py:55-68   GCNConv(cached=True, normalize=False) x2
* SAGE  /root/reference/tests/cpu/graphSAGE-sparse.py:65-78   SAGEConv(aggr=sum|mean) x2
* GIN   /root/reference/tests/cpu/gin-sparse.py:59-78   GINConv(MLP) x2 + BatchNorm + 2 Linear

With ``iSpLibPlugin.patch_pyg()`` active the matmul is the CUDA path; without it, it is the
stock (torch-op) matmul.  A layer can also be handed a ``spmm`` callable (e.g. a
``isplib_b200.dist.DistSpMM``) for the row-partitioned multi-GPU mode.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _matmul(adj_t, x, reduce):
    return sys.modules["torch_sparse"].matmul(adj_t, x, reduce)   # looked up per call: honours the patch


def _fused_available(x, spmm, adj_t=None) -> bool:
    """The fused-epilogue kernels run when the plugin is patched in (the CUDA path is the active
    matmul), the data is on the GPU and no external spmm callable (multi-GPU) was handed in."""
    if spmm is not None or not x.is_cuda:
        return False
    if getattr(adj_t, "is_partitioned", False) and os.environ.get("ISPLIB_B200_DIST_EPILOGUE", "1") == "0":
        return False        # opt-out: plain torch ops after the row-partitioned operator
    from . import iSpLibPlugin
    return iSpLibPlugin.is_patched() and iSpLibPlugin.fuse_epilogues


def _linear_padded(x, weight):
    """x @ weight.T written into rows padded to a multiple of 8 floats, returned as the [N, out]
    view: the SpMM that follows gathers these rows in place with 32-byte loads (out = 47 -> 48)
    instead of re-padding them on every call."""
    out = weight.size(0)
    pad = (-out) % 8
    if pad == 0 or out <= 8 or not x.is_cuda:
        return F.linear(x, weight)
    return F.linear(x, F.pad(weight, (0, 0, 0, pad))).narrow(1, 0, out)


class GCNConv(nn.Module):
    """out = A @ (x W) + b  (PyG GCNConv with normalize=False propagates AFTER the linear
    layer, at width out_channels).

    ``order``: 'linear_first' is PyG's order; 'aggregate_first' computes (A @ x) W, which is
    the same function but runs the SpMM at width in_channels and -- when x does not require
    grad, i.e. in the first layer -- needs no backward SpMM at all; 'auto' (default) picks
    aggregate_first iff in_channels < out_channels."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, order: str = "auto",
                 relu: bool = False):
        super().__init__()
        assert order in ("auto", "linear_first", "aggregate_first")
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.aggregate_first = (order == "aggregate_first") or (order == "auto" and in_channels < out_channels)
        self.relu = relu      # the activation that follows the layer (tests/cpu/gcn-sparse.py:64), fused when possible

    def forward(self, x, adj_t, spmm: Optional[Callable] = None):
        agg = (lambda t: spmm(t, "sum")) if spmm is not None else (lambda t: _matmul(adj_t, t, "sum"))
        if self.aggregate_first:
            out = F.linear(agg(x), self.lin.weight, self.bias)        # bias in the GEMM epilogue
            return F.relu(out) if self.relu else out
        if _fused_available(x, spmm, adj_t):
            # + bias and ReLU inside the SpMM's final store: two [N, K] passes less
            from . import fused_matmul
            if getattr(adj_t, "is_partitioned", False):      # the slice is staged into the peer-visible buffer anyway
                return fused_matmul(adj_t, F.linear(x, self.lin.weight), "sum", bias=self.bias, relu=self.relu)
            return fused_matmul(adj_t, _linear_padded(x, self.lin.weight), "sum", bias=self.bias, relu=self.relu)
        out = agg(self.lin(x))
        out = out if self.bias is None else out + self.bias
        return F.relu(out) if self.relu else out


class SAGEConv(nn.Module):
    """out = W_l * aggr_j(x_j) + W_r x_i ; the adjacency values are dropped (set_value(None))."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean"):
        super().__init__()
        self.aggr = aggr
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, adj_t, spmm: Optional[Callable] = None):
        if spmm is not None:
            agg = spmm(x, self.aggr)
        else:
            agg = _matmul(adj_t.set_value(None) if adj_t.has_value() else adj_t, x, self.aggr)
        # lin_l(agg) + lin_r(x) as ONE accumulate-GEMM: the root term (with lin_l's bias) is the
        # addmm's initial value, so no separate [N, out] add pass
        root = F.linear(x[: agg.size(0)], self.lin_r.weight, self.lin_l.bias)
        return torch.addmm(root, agg, self.lin_l.weight.t())


class GINConv(nn.Module):
    """out = MLP((1 + eps) x_i + sum_j x_j)."""

    def __init__(self, mlp: nn.Module, eps: float = 0.0):
        super().__init__()
        self.nn = mlp
        self.eps = eps

    def forward(self, x, adj_t, spmm: Optional[Callable] = None):
        adj = adj_t.set_value(None) if (spmm is None and adj_t.has_value()) else adj_t
        if _fused_available(x, spmm, adj) and adj.sparse_sizes()[0] == x.size(0):
            # (1 + eps) * x_i + sum_j x_j in the SpMM's final store (addend = x itself)
            from . import fused_matmul
            return self.nn(fused_matmul(adj, x, "sum", addend=x, addend_scale=1.0 + self.eps))
        agg = spmm(x, "sum") if spmm is not None else _matmul(adj, x, "sum")
        return self.nn((1.0 + self.eps) * x[: agg.size(0)] + agg)


class GCN(nn.Module):
    def __init__(self, in_channels, hidden, num_classes, dropout: float = 0.5, order: str = "auto"):
        super().__init__()
        self.conv1 = GCNConv(in_channels, hidden, order=order, relu=True)    # F.relu(conv1(...)), gcn-sparse.py:64
        self.conv2 = GCNConv(hidden, num_classes, order=order)
        self.dropout = dropout

    def forward(self, x, adj_t, spmm=None):
        x = self.conv1(x, adj_t, spmm)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.conv2(x, adj_t, spmm)
        return F.log_softmax(x, dim=1)


class GraphSAGE(nn.Module):
    def __init__(self, in_channels, hidden, num_classes, aggr="mean", dropout: float = 0.5):
        super().__init__()
        self.conv1 = SAGEConv(in_channels, hidden, aggr)
        self.conv2 = SAGEConv(hidden, num_classes, aggr)
        self.dropout = dropout

    def forward(self, x, adj_t, spmm=None):
        x = F.relu(self.conv1(x, adj_t, spmm))
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.conv2(x, adj_t, spmm)
        return F.log_softmax(x, dim=1)


class GIN(nn.Module):
    def __init__(self, in_channels, hidden, num_classes):
        super().__init__()
        mk = lambda i, o: nn.Sequential(nn.Linear(i, o), nn.ReLU(), nn.Linear(o, o))
        self.conv1 = GINConv(mk(in_channels, hidden))
        self.bn1 = nn.BatchNorm1d(hidden)
        self.conv2 = GINConv(mk(hidden, hidden))
        self.bn2 = nn.BatchNorm1d(hidden)
        self.fc1 = nn.Linear(hidden, hidden)
        self.fc2 = nn.Linear(hidden, num_classes)

    def forward(self, x, adj_t, spmm=None):
        x = self.bn1(F.relu(self.conv1(x, adj_t, spmm)))
        x = self.bn2(F.relu(self.conv2(x, adj_t, spmm)))
        x = F.dropout(F.relu(self.fc1(x)), p=0.5, training=self.training)
        return self.fc2(x)

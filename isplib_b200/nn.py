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

import sys
from typing import Callable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _matmul(adj_t, x, reduce):
    return sys.modules["torch_sparse"].matmul(adj_t, x, reduce)   # looked up per call: honours the patch


class GCNConv(nn.Module):
    """out = A @ (x W) + b  (PyG GCNConv with normalize=False propagates AFTER the linear
    layer, at width out_channels).

    ``order``: 'linear_first' is PyG's order; 'aggregate_first' computes (A @ x) W, which is
    the same function but runs the SpMM at width in_channels and -- when x does not require
    grad, i.e. in the first layer -- needs no backward SpMM at all; 'auto' (default) picks
    aggregate_first iff in_channels < out_channels."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, order: str = "auto"):
        super().__init__()
        assert order in ("auto", "linear_first", "aggregate_first")
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.aggregate_first = (order == "aggregate_first") or (order == "auto" and in_channels < out_channels)

    def forward(self, x, adj_t, spmm: Optional[Callable] = None):
        agg = (lambda t: spmm(t, "sum")) if spmm is not None else (lambda t: _matmul(adj_t, t, "sum"))
        if self.aggregate_first:
            out = self.lin(agg(x))
        else:
            out = agg(self.lin(x))
        return out if self.bias is None else out + self.bias


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
        return self.lin_l(agg) + self.lin_r(x[: agg.size(0)])


class GINConv(nn.Module):
    """out = MLP((1 + eps) x_i + sum_j x_j)."""

    def __init__(self, mlp: nn.Module, eps: float = 0.0):
        super().__init__()
        self.nn = mlp
        self.eps = eps

    def forward(self, x, adj_t, spmm: Optional[Callable] = None):
        if spmm is not None:
            agg = spmm(x, "sum")
        else:
            agg = _matmul(adj_t.set_value(None) if adj_t.has_value() else adj_t, x, "sum")
        return self.nn((1.0 + self.eps) * x[: agg.size(0)] + agg)


class GCN(nn.Module):
    def __init__(self, in_channels, hidden, num_classes, dropout: float = 0.5, order: str = "auto"):
        super().__init__()
        self.conv1 = GCNConv(in_channels, hidden, order=order)
        self.conv2 = GCNConv(hidden, num_classes, order=order)
        self.dropout = dropout

    def forward(self, x, adj_t, spmm=None):
        x = F.relu(self.conv1(x, adj_t, spmm))
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

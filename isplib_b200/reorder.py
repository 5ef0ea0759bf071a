"""Locality reordering of the graph around the hot path.

The reference's "autotuner" exists to pick a CPU-friendly feature width
(/root/reference/autotuner/findbestk.py) and its loaders pad features to multiples of 16
(/root/reference/tests/cpu/dataset_loader.py:145-160).  On B200 the K side is handled inside
the op layer (rows padded to 16 bytes, on-device variant selection); what is left to the data
side is the ORDER of the nodes: the SpMM is bound by gathers of X rows, and a node order that
keeps a row's neighbours close together turns L2 misses into L2/L1 hits.

    perm = reverse_cuthill_mckee(adj)            # or degree_order(adj), or any permutation
    adj_p = permute(adj, perm)                   # P A P^T
    out_p = torch_sparse.matmul(adj_p, x[perm])  # == (A x)[perm]

Symmetric permutations commute with the SpMM, so training is unchanged up to the fixed
relabelling of nodes (features, labels and masks are indexed with the same ``perm``).
"""
from __future__ import annotations

import sys

import numpy as np
import torch


def _st():
    import isplib_b200  # noqa: F401
    return sys.modules["torch_sparse"].SparseTensor


def degree_order(adj, descending: bool = True) -> torch.Tensor:
    """Nodes sorted by degree: hubs first -- the rows every other row gathers share cache lines."""
    deg = adj.storage.rowcount()
    return torch.argsort(deg, descending=descending, stable=True)


def reverse_cuthill_mckee(adj) -> torch.Tensor:
    """Bandwidth-reducing order (scipy's RCM on the symmetrised pattern; host side, one-time)."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee as rcm
    rowptr, col, _ = adj.csr()
    m, n = adj.sparse_sizes()
    assert m == n, "symmetric reordering needs a square adjacency"
    a = sp.csr_matrix((np.ones(col.numel(), dtype=np.int8), col.cpu().numpy(), rowptr.cpu().numpy()), shape=(m, n))
    perm = rcm((a + a.T).tocsr(), symmetric_mode=True)
    return torch.from_numpy(np.ascontiguousarray(perm).astype(np.int64)).to(col.device)


def bfs_order(adj, start=None) -> torch.Tensor:
    """Breadth-first level order (Cuthill-McKee without the per-level degree sort being sequential):
    nodes sorted by (BFS level from a max-degree seed, degree ascending inside a level), computed
    WHERE THE GRAPH LIVES with data-parallel frontier expansions (one masked pass over the edge list
    per level; Reddit-shaped graphs have < 10 levels) -- the device-side alternative to the host-side
    scipy RCM above (1.6 s on a 30 M-entry graph, profiles/r1_reorder_demo.txt).  Unreached
    components are seeded again from their highest-degree node."""
    rowptr, col, _ = adj.csr()
    m, n = adj.sparse_sizes()
    assert m == n, "symmetric reordering needs a square adjacency"
    dev = col.device
    deg = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(m, device=dev), deg)
    level = torch.full((m,), -1, dtype=torch.int64, device=dev)
    cur = 0
    while True:
        unreached = level < 0
        if not bool(unreached.any()):
            break
        if start is not None and cur == 0:
            seed = int(start)
        else:
            seed = int(torch.argmax(torch.where(unreached, deg, torch.full_like(deg, -1))))
        frontier = torch.zeros(m, dtype=torch.bool, device=dev)
        frontier[seed] = True
        level[seed] = cur
        while True:
            # neighbours (both directions: the pattern is treated as undirected) of the frontier
            hit_out = frontier[row] & (level[col] < 0)
            hit_in = frontier[col] & (level[row] < 0)
            nxt = torch.zeros(m, dtype=torch.bool, device=dev)
            nxt[col[hit_out]] = True
            nxt[row[hit_in]] = True
            if not bool(nxt.any()):
                break
            cur += 1
            level[nxt] = cur
            frontier = nxt
        cur += 1
    key = level * (int(deg.max()) + 1 if m else 1) + deg
    return torch.argsort(key, stable=True)


def permute(adj, perm: torch.Tensor):
    """P A P^T: new node i is old node perm[i].  Built through the device COO->CSR path when the
    graph lives on a GPU, with torch ops otherwise."""
    row, col, val = adj.coo()
    m, n = adj.sparse_sizes()
    assert m == n and perm.numel() == m
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(m, device=perm.device, dtype=perm.dtype)
    new_row, new_col = inv[row], inv[col]
    if row.is_cuda:
        from . import io
        return io.from_edge_index(torch.stack([new_row, new_col]), val, m, n, device=row.device)
    return _st()(row=new_row, col=new_col, value=val, sparse_sizes=(m, n))

"""Graph ingest formats on either side of the hot path.

* ``from_edge_index`` -- a PyG ``edge_index`` ([2, E] int64) + optional edge weights ->
  ``SparseTensor`` in CSR, built ON THE DEVICE by ``isplib_b200_coo_to_csr`` (stable radix sort
  by (row, col)) instead of torch_sparse's host-side argsort that every loader of the reference
  goes through (/root/reference/tests/cpu/dataset_loader.py:10, ``T.ToSparseTensor()``).
* ``read_mtx`` / ``write_mtx`` -- Matrix Market coordinate files, the format the reference's
  autotuner workflow is built around (/root/reference/autotuner/findbestk.py:36,
  README.md:144-169).  Parsing is scipy's; the CSR build is the device path above.
"""
from __future__ import annotations

import sys
from typing import Optional

import torch


def _sparse_tensor_cls():
    import isplib_b200  # noqa: F401  (installs the torch_sparse stand-in when the package is absent)
    return sys.modules["torch_sparse"].SparseTensor


def from_edge_index(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None,
                    num_rows: Optional[int] = None, num_cols: Optional[int] = None, device="cuda"):
    """adj with adj[row, col] = attr for every (row, col) = edge_index[:, e].  For a PyG
    message-passing graph pass ``edge_index.flip(0)`` (adj_t = transposed adjacency)."""
    from . import capi
    dev = torch.device(device)
    row = edge_index[0].to(dev)
    col = edge_index[1].to(dev)
    m = int(num_rows) if num_rows is not None else (int(row.max()) + 1 if row.numel() else 0)
    n = int(num_cols) if num_cols is not None else (int(col.max()) + 1 if col.numel() else 0)
    r32 = row.to(torch.int32).contiguous()
    c32 = col.to(torch.int32).contiguous()
    val = None if edge_attr is None else edge_attr.to(dev, torch.float32).contiguous()
    rowptr, col_s, val_s, _ = capi.coo_to_csr(r32, c32, val, m, n)
    ST = _sparse_tensor_cls()
    return ST(rowptr=rowptr.to(torch.int64), col=col_s.to(torch.int64), value=val_s, sparse_sizes=(m, n), is_sorted=True)


def read_mtx(path: str, device="cuda", pattern_as_none: bool = True):
    """Matrix Market coordinate file -> SparseTensor on ``device`` (CSR built on the device)."""
    import scipy.io
    coo = scipy.io.mmread(path, spmatrix=False).tocoo() if "spmatrix" in scipy.io.mmread.__code__.co_varnames else scipy.io.mmread(path).tocoo()
    ei = torch.stack([torch.from_numpy(coo.row.astype("int64")), torch.from_numpy(coo.col.astype("int64"))])
    with open(path, "rb") as f:
        header = f.readline().decode("ascii", "replace").lower()
    attr = None if (pattern_as_none and "pattern" in header) else torch.from_numpy(coo.data.astype("float32"))
    return from_edge_index(ei, attr, coo.shape[0], coo.shape[1], device)


def write_mtx(adj, path: str) -> None:
    """SparseTensor -> Matrix Market (the export step of README.md:144-169, without
    fast_matrix_market).  Duplicate entries are summed, as coalesce() does there."""
    import scipy.io
    import scipy.sparse
    row, col, val = adj.coo()
    m, n = adj.sparse_sizes()
    v = torch.ones(col.numel()) if val is None else val.detach().cpu()
    coo = scipy.sparse.coo_matrix((v.numpy(), (row.cpu().numpy(), col.cpu().numpy())), shape=(m, n))
    coo.sum_duplicates()
    scipy.io.mmwrite(path, coo)

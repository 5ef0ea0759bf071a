"""Partitioned ingest: from an edge list that is itself spread over the ranks (a shard of ``edge_index`` per
process, or every process reading its own byte range of one Matrix Market file) to the row-partitioned
operator of ``isplib_b200.dist`` -- without any rank ever holding the whole graph.

The multi-GPU counterpart of ``isplib_b200.io`` (the reference is single-process
and loads everything through one host-side ``T.ToSparseTensor()``, /root/reference/tests/cpu/dataset_loader.py:10,
or one Matrix Market read, /root/reference/autotuner/findbestk.py:36):

1. the degree histogram of the local edges is summed over the ranks (M integers) -> row ranges balanced by
   stored entries (``nnz_balanced_bounds``, the same cut the whole-graph constructor makes);
2. every edge travels to the owner of its row (``exchange_by_owner``: one point-to-point exchange);
3. each rank builds the CSR of its own rows on ITS device (``isplib_b200_coo_to_csr``, stable radix sort by
   (row, col) -- the kernel ``io.from_edge_index`` uses);
4. ``DistSpMM.from_local_rows`` assembles the operator (one all-gather of the degrees; the transposed
   partition for the backward by one more exchange).

Entries with equal (row, col) keep the order "source rank, then position in that rank's shard", which is the
file / edge_index order when the shards are contiguous pieces in rank order.
"""
from __future__ import annotations

import io as _io
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from .dist import PartitionedAdj, even_bounds, exchange_by_owner, nnz_balanced_bounds


def _cuda_csr_builder(row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], m: int, n: int):
    """(rowptr int64 [m + 1], col int64, val) of the rank's own rows, on the device the edges are on."""
    from . import capi
    if not row.is_cuda:
        raise RuntimeError("isplib_b200.dist_io: the CSR of a row block is built by the CUDA library "
                           "(isplib_b200_coo_to_csr); move the edges to the rank's GPU (device=...)")
    rowptr, col_s, val_s, _ = capi.coo_to_csr(row.to(torch.int32).contiguous(), col.to(torch.int32).contiguous(),
                                              None if val is None else val.to(torch.float32).contiguous(), m, n)
    return rowptr.to(torch.int64), col_s.to(torch.int64), val_s


def _raise_together(err: Optional[Exception], group, device) -> None:
    """Every rank raises if ANY rank holds an error: an exception on one rank alone would leave the others
    waiting in the next collective forever."""
    flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if err is not None:
        raise err
    if int(flag):
        raise RuntimeError("isplib_b200.dist_io: another rank of the group rejected its shard of the input "
                           "(its own exception names the cause)")


def partition_edges(row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], num_rows: int, num_cols: int,
                    group=None, device=None, balance: str = "nnz", csr_builder: Optional[Callable] = None,
                    **dist_kw) -> PartitionedAdj:
    """This rank's shard of the edges (adj[row[e], col[e]] = val[e], global ids; any rows, any order) ->
    the row-partitioned adjacency over ``group``.  Collective: every rank of the group calls it."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device(device) if device is not None else row.device
    row = row.to(dev).to(torch.int64)
    col = col.to(dev).to(torch.int64)
    val = None if val is None else val.to(dev).to(torch.float32)
    m, n = int(num_rows), int(num_cols)
    if balance not in ("nnz", "rows"):
        raise ValueError(f"balance must be 'nnz' or 'rows', got {balance!r}")   # same argument on every rank
    err = None
    if row.numel() != col.numel() or (val is not None and val.numel() != row.numel()):
        err = ValueError("isplib_b200.dist_io: row, col and val of a shard must have the same length")
    elif row.numel() and (int(row.min()) < 0 or int(row.max()) >= m or int(col.min()) < 0 or int(col.max()) >= n):
        err = ValueError("isplib_b200.dist_io: edge endpoint outside [0, num_rows) x [0, num_cols)")
    _raise_together(err, group, dev)
    if balance == "nnz":
        deg = torch.bincount(row, minlength=m) if row.numel() else torch.zeros(m, dtype=torch.int64, device=dev)
        dist.all_reduce(deg, group=group)
        rowptr = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(deg, 0)
        bounds = nnz_balanced_bounds(rowptr, world)
    else:
        bounds = even_bounds(m, world)
    cuts = torch.as_tensor(bounds, dtype=torch.int64, device=dev)
    owner = torch.bucketize(row, cuts[1:-1], right=True)
    got = exchange_by_owner([row, col] + ([val] if val is not None else []), owner, group)
    r0, r1 = bounds[rank], bounds[rank + 1]
    build = csr_builder or _cuda_csr_builder
    rowptr_l, col_l, val_l = build(got[0] - r0, got[1], got[2] if val is not None else None, r1 - r0, n)
    return PartitionedAdj(rowptr_l, col_l, val_l, n, local_rows=True, group=group, device=dev, **dist_kw)


# ------------------------------------------------------------------------------------------------------------
# Matrix Market, one byte range per rank
# ------------------------------------------------------------------------------------------------------------
def mtx_header(path: str):
    """(num_rows, num_cols, num_entries, field, symmetry, data_offset) of a coordinate Matrix Market file."""
    with open(path, "rb") as f:
        banner = f.readline().decode("ascii", "replace").strip().lower().split()
        if len(banner) < 5 or banner[0] != "%%matrixmarket" or banner[1] != "matrix":
            raise ValueError(f"{path}: not a Matrix Market file")
        fmt, field, symmetry = banner[2], banner[3], banner[4]
        if fmt != "coordinate":
            raise ValueError(f"{path}: only the coordinate (sparse) format holds a graph, got {fmt!r}")
        if field not in ("real", "integer", "pattern", "double"):
            raise NotImplementedError(f"{path}: field {field!r} (fp32 real values or pattern only)")
        if symmetry not in ("general", "symmetric"):
            raise NotImplementedError(f"{path}: symmetry {symmetry!r}")
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no size line")
            s = line.strip()
            if s and not s.startswith(b"%"):
                m, n, nnz = (int(t) for t in s.split()[:3])
                return m, n, nnz, field, symmetry, f.tell()


def _line_start(f, pos: int, lo: int, hi: int) -> int:
    """First line start at or after byte `pos` (a line starts at `lo` or right after a newline)."""
    if pos <= lo:
        return lo
    if pos >= hi:
        return hi
    f.seek(pos - 1)
    while True:
        buf = f.read(1 << 16)
        if not buf:
            return hi
        k = buf.find(b"\n")
        if k >= 0:
            return min(hi, f.tell() - len(buf) + k + 1)


def read_mtx_shard(path: str, rank: int, world: int):
    """(row, col, val | None, num_rows, num_cols): the entries whose LINES start inside this rank's byte range
    of the data section -- every line is read by exactly one rank, whatever the line lengths.  0-based int64
    ids; a symmetric file's off-diagonal entries are mirrored."""
    import numpy as np
    m, n, _, field, symmetry, data0 = mtx_header(path)
    size = os.path.getsize(path)
    span = size - data0
    with open(path, "rb") as f:
        a = _line_start(f, data0 + span * rank // world, data0, size)
        b = _line_start(f, data0 + span * (rank + 1) // world, data0, size)
        f.seek(a)
        chunk = f.read(b - a)
    ncol = 2 if field == "pattern" else 3
    if chunk.strip():
        try:
            import pandas as pd
            arr = pd.read_csv(_io.BytesIO(chunk), sep=r"\s+", header=None, comment="%", dtype=np.float64,
                              engine="c").to_numpy()
        except ImportError:                                   # pragma: no cover
            arr = np.loadtxt(_io.BytesIO(chunk), comments="%", ndmin=2, dtype=np.float64)
    else:
        arr = np.zeros((0, ncol), dtype=np.float64)
    if arr.shape[0] and arr.shape[1] < ncol:
        raise ValueError(f"{path}: expected {ncol} columns per entry, found {arr.shape[1]}")
    row = arr[:, 0].astype(np.int64) - 1
    col = arr[:, 1].astype(np.int64) - 1
    val = None if field == "pattern" else arr[:, 2].astype(np.float32)
    if symmetry == "symmetric":
        off = row != col
        row, col = np.concatenate([row, col[off]]), np.concatenate([col, row[off]])
        if val is not None:
            val = np.concatenate([val, val[off]])
    return (torch.from_numpy(row), torch.from_numpy(col), None if val is None else torch.from_numpy(val), m, n)


def read_mtx_partitioned(path: str, group=None, device="cuda", pattern_as_none: bool = True, **kw) -> PartitionedAdj:
    """Every rank parses its own byte range of one Matrix Market file and the edges are routed to their row
    owners: ``io.read_mtx`` + ``iSpLibPlugin.partition`` without the whole graph in any one process."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    err, shard = None, None
    try:
        shard = read_mtx_shard(path, rank, world)
    except Exception as e:          # a malformed line in ONE byte range must stop every rank, not hang the rest
        err = e
    _raise_together(err, group, torch.device(device))
    row, col, val, m, n = shard
    if val is None and not pattern_as_none:
        val = torch.ones(row.numel(), dtype=torch.float32)
    return partition_edges(row, col, val, m, n, group=group, device=device, **kw)

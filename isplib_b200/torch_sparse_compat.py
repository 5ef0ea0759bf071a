"""Minimal stand-in for the ``torch_sparse`` package.

The reference imports ``torch_sparse`` unconditionally
(/root/reference/isplib/__init__.py:6-8) and patches ``torch_sparse.matmul``.  That
package is absent from this image, so when it cannot be imported ``isplib_b200``
installs this module as ``sys.modules['torch_sparse']``.  It provides exactly what the
reference touches:

* ``SparseTensor(row=, col=, value=, sparse_sizes=, rowptr=)``, ``.csr()``,
  ``.storage._row/_rowcount/_csr2csc/_colptr`` and the ``row()/rowcount()/csr2csc()/
  colptr()`` accessors (isplib/__init__.py:49,58-73), ``set_value``, ``t()``,
  ``to()/cuda()``, ``from_edge_index``, ``to_torch_sparse_coo_tensor`` (README.md:153);
* a module-level ``matmul(src, other, reduce)`` -- the stock, UNPATCHED behaviour,
  written with plain torch ops.  It is what runs when the plugin is *not* active (the
  reference's "pt1" mode); once ``iSpLibPlugin.patch_pyg()`` has run, ``matmul`` is the
  CUDA path and this function is not involved.

When the real torch_sparse is installed it is used instead and this file is inert.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

__version__ = "0.0.0+isplib_b200.compat"


class SparseStorage:
    def __init__(self, row, rowptr, col, value, sparse_sizes, is_sorted=False):
        assert col is not None and (row is not None or rowptr is not None)
        M, N = sparse_sizes
        if row is None:
            deg = rowptr[1:] - rowptr[:-1]
            row = torch.repeat_interleave(torch.arange(M, device=col.device, dtype=col.dtype), deg)
            is_sorted = True  # a rowptr is only meaningful for row-sorted entries
        if M is None:
            M = int(row.max()) + 1 if row.numel() else 0
        if N is None:
            N = int(col.max()) + 1 if col.numel() else 0
        if not is_sorted and row.numel() > 1:
            key = row * N + col
            if not bool((key[1:] >= key[:-1]).all()):
                perm = torch.argsort(key, stable=True)
                row, col = row[perm], col[perm]
                value = None if value is None else value[perm]
                rowptr = None
        self._sparse_sizes = (int(M), int(N))
        self._row = row
        self._rowptr = rowptr
        self._col = col
        self._value = value
        self._rowcount = None
        self._colptr = None
        self._colcount = None
        self._csr2csc = None
        self._csc2csr = None

    # -- accessors the plugin calls (isplib/__init__.py:69-73) --------------------
    def sparse_sizes(self): return self._sparse_sizes
    def col(self): return self._col
    def value(self): return self._value
    def has_value(self): return self._value is not None

    def row(self):
        return self._row

    def rowptr(self):
        if self._rowptr is None:
            M = self._sparse_sizes[0]
            counts = torch.bincount(self._row, minlength=M) if self._row.numel() else \
                torch.zeros(M, dtype=torch.long, device=self._col.device)
            rp = torch.zeros(M + 1, dtype=self._col.dtype, device=self._col.device)
            rp[1:] = torch.cumsum(counts, 0)
            self._rowptr = rp
        return self._rowptr

    def rowcount(self):
        if self._rowcount is None:
            rp = self.rowptr()
            self._rowcount = rp[1:] - rp[:-1]
        return self._rowcount

    def csr2csc(self):
        if self._csr2csc is None:
            M = self._sparse_sizes[0]
            self._csr2csc = torch.argsort(self._col * M + self._row, stable=True)
        return self._csr2csc

    def colptr(self):
        if self._colptr is None:
            N = self._sparse_sizes[1]
            counts = torch.bincount(self._col, minlength=N) if self._col.numel() else \
                torch.zeros(N, dtype=torch.long, device=self._col.device)
            cp = torch.zeros(N + 1, dtype=self._col.dtype, device=self._col.device)
            cp[1:] = torch.cumsum(counts, 0)
            self._colptr = cp
        return self._colptr

    def set_value(self, value):
        st = SparseStorage.__new__(SparseStorage)
        st.__dict__.update(self.__dict__)
        st._value = value
        return st

    def to(self, *args, **kwargs):
        st = SparseStorage.__new__(SparseStorage)
        st.__dict__.update(self.__dict__)
        for name in ("_row", "_rowptr", "_col", "_rowcount", "_colptr", "_colcount", "_csr2csc", "_csc2csr"):
            t = getattr(self, name)
            if t is not None:
                dev_only = {k: v for k, v in kwargs.items() if k in ("device", "non_blocking")}
                devs = [a for a in args if isinstance(a, (str, torch.device))]
                setattr(st, name, t.to(*devs, **dev_only))
        if self._value is not None:
            st._value = self._value.to(*args, **kwargs)
        return st


class SparseTensor:
    def __init__(self, row: Optional[torch.Tensor] = None, rowptr: Optional[torch.Tensor] = None,
                 col: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                 sparse_sizes: Optional[Tuple[Optional[int], Optional[int]]] = None,
                 is_sorted: bool = False, trust_data: bool = False):
        if sparse_sizes is None:
            sparse_sizes = (None, None)
        self.storage = SparseStorage(row, rowptr, col, value, sparse_sizes, is_sorted)

    @classmethod
    def from_storage(cls, storage):
        self = cls.__new__(cls)
        self.storage = storage
        return self

    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None, is_sorted=False, trust_data=False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes,
                   is_sorted=is_sorted)

    # -- what the plugin uses ------------------------------------------------------
    def csr(self):
        return self.storage.rowptr(), self.storage.col(), self.storage.value()

    def coo(self):
        return self.storage.row(), self.storage.col(), self.storage.value()

    def sparse_sizes(self): return self.storage.sparse_sizes()
    def sparse_size(self, dim): return self.storage.sparse_sizes()[dim]
    def size(self, dim): return self.storage.sparse_sizes()[dim]
    def sizes(self): return list(self.storage.sparse_sizes())
    def nnz(self): return self.storage.col().numel()
    def has_value(self): return self.storage.has_value()
    @property
    def device(self): return self.storage.col().device
    def is_cuda(self): return self.storage.col().is_cuda

    def set_value(self, value, layout=None):
        return SparseTensor.from_storage(self.storage.set_value(value))

    def fill_cache_(self):
        self.storage.rowptr(); self.storage.rowcount(); self.storage.csr2csc(); self.storage.colptr()
        return self

    def to(self, *args, **kwargs):
        return SparseTensor.from_storage(self.storage.to(*args, **kwargs))

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def cpu(self):
        return self.to("cpu")

    def t(self):
        row, col, value = self.coo()
        M, N = self.sparse_sizes()
        return SparseTensor(row=col, col=row, value=value, sparse_sizes=(N, M))

    def to_torch_sparse_coo_tensor(self):
        row, col, value = self.coo()
        if value is None:
            value = torch.ones(col.numel(), device=col.device)
        return torch.sparse_coo_tensor(torch.stack([row, col]), value, self.sparse_sizes())

    def to_dense(self):
        return self.to_torch_sparse_coo_tensor().to_dense()

    def matmul(self, other, reduce: str = "sum"):
        import torch_sparse  # this module (or the real package): honours an active patch
        return torch_sparse.matmul(self, other, reduce)

    __matmul__ = matmul

    def __repr__(self):
        return f"SparseTensor(sizes={self.sparse_sizes()}, nnz={self.nnz()}, has_value={self.has_value()})"


def _stock_matmul(src: SparseTensor, other: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """Stock torch_sparse.matmul semantics with plain torch ops (any device)."""
    row, col, value = src.coo()
    M = src.size(0)
    msg = other.index_select(0, col)
    if value is not None:
        msg = msg * value.to(other.dtype).unsqueeze(-1)
    out = torch.zeros((M, other.size(1)), dtype=other.dtype, device=other.device)
    if reduce in ("sum", "add"):
        return out.index_add_(0, row, msg)
    if reduce == "mean":
        out.index_add_(0, row, msg)
        deg = src.storage.rowcount().clamp(min=1).to(other.dtype)
        return out / deg.unsqueeze(-1)
    if reduce in ("max", "min"):
        idx = row.unsqueeze(-1).expand_as(msg)
        return out.scatter_reduce_(0, idx, msg, "amax" if reduce == "max" else "amin", include_self=False)
    raise ValueError(f"unsupported reduce: {reduce!r}")


def matmul(src, other, reduce: str = "sum"):
    if isinstance(src, SparseTensor) and isinstance(other, torch.Tensor):
        return _stock_matmul(src, other, reduce)
    raise ValueError("torch_sparse compat: only SparseTensor @ dense Tensor is provided")

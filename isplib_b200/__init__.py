"""isplib_b200 -- B200-native drop-in for iSpLib's one hot path.

Same Python surface as /root/reference/isplib/__init__.py:
``iSpLibPlugin.patch_pyg()`` / ``unpatch_pyg()``, the ``@isplib_autotune`` decorator and
``torch_sparse.matmul(adj_t, X, reduce)`` diverted to ``torch.ops.isplib.fusedmm_spmm*``
-- but the ops are hand-written sm_100a CUDA kernels behind the C ABI of
``include/isplib_b200.h``.  There is no CPU path: importing this package without the
built extension raises ImportError, and CPU tensors are rejected by the ops.

``import isplib`` (the thin alias package at the repo root) re-exports everything here,
so ``from isplib import *`` at the top of an existing PyG script keeps working.
"""
from __future__ import annotations

import functools
import importlib.machinery
import os
import os.path as osp
import sys

import torch

__version__ = "0.2.0+b200.r2"

# --- torch_sparse: the real package if present, the compat shim otherwise -------------
try:  # /root/reference/isplib/__init__.py:6-8 imports it unconditionally
    import torch_sparse  # type: ignore
except Exception:  # absent in this image
    from . import torch_sparse_compat as torch_sparse
    sys.modules.setdefault("torch_sparse", torch_sparse)

from torch_sparse import SparseTensor, matmul  # noqa: E402  (re-exported like the reference)

try:  # isplib/__init__.py:10-13
    import torch_geometric.typing  # type: ignore  # noqa: F401
except Exception:
    pass

# --- extension loader: mirrors isplib/__init__.py:18-28, CUDA only ---------------------
_PKG_DIR = osp.dirname(osp.abspath(__file__))


def _load_extension() -> str:
    spec = importlib.machinery.PathFinder().find_spec("_fusedmm_cuda", [_PKG_DIR])
    if spec is None or spec.origin is None:
        raise ImportError(
            f"Could not find module '_fusedmm_cuda' in {_PKG_DIR}. Build it with "
            "`make -C isplib_b200/csrc all ops` (or __graft_entry__.build()). "
            "isplib_b200 has no CPU implementation to fall back to.")
    torch.ops.load_library(spec.origin)
    return spec.origin


EXTENSION_PATH = None
if os.environ.get("ISPLIB_B200_SKIP_EXTENSION", "0") != "1":
    EXTENSION_PATH = _load_extension()


def _is_sparse_tensor(obj) -> bool:
    return hasattr(obj, "csr") and hasattr(obj, "storage") and not isinstance(obj, torch.Tensor)


class iSpLibPlugin:
    """Same class-level surface as the reference (isplib/__init__.py:34-40).  The caches
    are kept for compatibility; the derived per-graph data (int32 CSR, segment plan, CSC
    view, permuted values, tuned variant) now lives in the C++ op layer, keyed by the
    graph's storage and dropped when the graph's tensors die (the reference keys by raw
    data_ptr and never evicts, isplib/__init__.py:50)."""
    backup = []
    value_cache = {}
    cache = {}
    row_cache = {}
    is_cached = False
    value_cached = False
    # 'reference': max/min leave lowest()/max() in rows without entries, as
    # csrc/fusedmm.cpp:147-150 does; 'zero': torch_sparse's convention (what PyG's aggr='max'
    # expects for isolated nodes).  arg_out is the nnz sentinel in both.
    empty_row_mode = "reference"
    # isplib_b200.nn layers fuse their epilogues (bias, ReLU, GIN's self term) into the SpMM while
    # the plugin is patched in; False keeps them as separate torch ops (A/B and debugging)
    fuse_epilogues = True

    @classmethod
    def set_empty_row_mode(cls, mode: str):
        if mode not in ("reference", "zero"):
            raise ValueError("empty_row_mode must be 'reference' or 'zero'")
        cls.empty_row_mode = mode
        if mode == "zero":
            os.environ["ISPLIB_B200_EMPTY_ROWS"] = "zero"
        else:
            os.environ.pop("ISPLIB_B200_EMPTY_ROWS", None)

    @staticmethod
    def spmm(src, other, reduce: str = "sum"):
        """The patched ``torch_sparse.matmul``: body of ``spmm_autotuned``
        (isplib/__init__.py:48-157) on the CUDA ops."""
        if getattr(src, "is_partitioned", False):
            # a row-partitioned adjacency (isplib_b200.dist.partition / iSpLibPlugin.partition): `other` is
            # this rank's slice of X; the multi-GPU operator (fused gather + SpMM kernel, autograd) runs
            return src.matmul(other, reduce)
        if not _is_sparse_tensor(src):
            # torch.sparse.mm is patched too (isplib/__init__.py:178); genuine torch sparse
            # tensors go to the original function instead of crashing on src.csr()
            for orig in reversed(iSpLibPlugin.backup[1::2]):      # the saved torch.sparse.mm's
                if orig is not iSpLibPlugin.spmm:
                    return orig(src, other)
            raise TypeError("isplib: expected a torch_sparse.SparseTensor")
        rowptr, col, value = src.csr()
        if value is not None:
            value = value.to(other.dtype)                     # isplib/__init__.py:63-64
        # value None stays None: the C ABI treats a null value pointer as all-ones, so
        # the 4*nnz-byte ones vector of isplib/__init__.py:51-57 is never materialised.
        st = src.storage
        row, rowcount, csr2csc, colptr = st._row, st._rowcount, st._csr2csc, st._colptr
        ops = torch.ops.isplib
        if reduce in ("sum", "add"):
            return ops.fusedmm_spmm(row, rowptr, col, value, colptr, csr2csc, other, None, None)
        if reduce == "mean":
            return ops.fusedmm_spmm_mean(row, rowptr, col, value, rowcount, colptr, csr2csc, other, None, None)
        if reduce == "max":
            # the op returns (out, arg_out) like the reference's; matmul returns `out`
            # (the torch_sparse contract; the reference leaks the tuple, isplib/__init__.py:143)
            return ops.fusedmm_spmm_max(rowptr, col, value, other)[0]
        if reduce == "min":
            return ops.fusedmm_spmm_min(rowptr, col, value, other)[0]
        raise ValueError(f"isplib: unsupported reduce {reduce!r} (sum, add, mean, max, min)")

    dist_group = None      # process group of the row-partitioned mode (patch_pyg(group=...))
    _group_backup = []     # the groups of the enclosing patch scopes (the patch is a stack, like `backup`)

    @classmethod
    def partition(cls, adj_t, device=None, **kw):
        """Row-partition `adj_t` over the ranks of the group given to ``patch_pyg(group=...)`` (default:
        the world group).  The result goes wherever the SparseTensor went: ``matmul(padj, x_slice, reduce)``.
        ``local_rows=True``: `adj_t` is only this rank's block of rows (rank-local ingest, see dist.partition)."""
        from .dist import partition
        return partition(adj_t, group=cls.dist_group, device=device, **kw)

    @classmethod
    def patch_pyg(cls, group=None):
        """``group``: a torch.distributed process group -- one process per GPU -- for the row-partitioned
        multi-GPU mode: adjacencies made with ``iSpLibPlugin.partition(adj_t)`` are then multiplied
        across its ranks by the same patched ``torch_sparse.matmul`` (new; the reference is single-process)."""
        global matmul
        cls._group_backup.append(cls.dist_group)
        cls.dist_group = group
        try:  # isplib/__init__.py:159-171
            import torch_geometric.typing as tgt  # type: ignore
            cls.cache["WITH_PT2"] = getattr(tgt, "WITH_PT2", None)
            cls.cache["WITH_PT20"] = getattr(tgt, "WITH_PT20", None)
            tgt.WITH_PT2 = False
            tgt.WITH_PT20 = False
        except Exception:
            pass
        ts = sys.modules["torch_sparse"]
        cls.backup.append(ts.matmul)                          # isplib/__init__.py:173-174
        cls.backup.append(torch.sparse.mm)
        ts.matmul = cls.spmm                                  # isplib/__init__.py:177-178
        torch.sparse.mm = cls.spmm
        matmul = cls.spmm

    @classmethod
    def unpatch_pyg(cls):
        global matmul
        if len(cls.backup) > 0:                               # isplib/__init__.py:190-195
            torch.sparse.mm = cls.backup.pop()
            ts = sys.modules["torch_sparse"]
            ts.matmul = cls.backup.pop()
            matmul = ts.matmul
            cls.dist_group = cls._group_backup.pop() if cls._group_backup else None
            try:
                import torch_geometric.typing as tgt  # type: ignore
                if cls.cache.get("WITH_PT2") is not None:
                    tgt.WITH_PT2 = cls.cache["WITH_PT2"]
                if cls.cache.get("WITH_PT20") is not None:
                    tgt.WITH_PT20 = cls.cache["WITH_PT20"]
            except Exception:
                pass

    @classmethod
    def is_patched(cls) -> bool:
        return len(cls.backup) > 0


def fused_matmul(src, other, reduce: str = "sum", bias=None, addend=None, addend_scale: float = 1.0,
                 relu: bool = False):
    """``relu?(matmul(src, other, reduce) + addend_scale * addend + bias)`` in ONE kernel: what the
    reference's callers do to the SpMM result in separate [M, K] passes -- GCNConv's bias + ReLU
    (/root/reference/tests/cpu/gcn-sparse.py:61-68), GINConv's ``(1 + eps) * x_i + aggr``
    (gin-sparse.py:73-78; pass ``addend=other``) -- fused into the kernel's final store
    (``isplib_b200_spmm_csr_fused``).  sum / add / mean; differentiable w.r.t. other, value, bias,
    addend.  CUDA tensors only, like every op of this package."""
    if getattr(src, "is_partitioned", False):      # row-partitioned over a process group: same epilogue, last launch
        return src.matmul(other, reduce, bias=bias, addend=addend, addend_scale=addend_scale, relu=relu)
    if not _is_sparse_tensor(src):
        raise TypeError("isplib: expected a torch_sparse.SparseTensor")
    rowptr, col, value = src.csr()
    if value is not None:
        value = value.to(other.dtype)
    return torch.ops.isplib.fusedmm_spmm_fused(rowptr, col, value, other, reduce, bias, addend,
                                               float(addend_scale), bool(relu))


def pad_features(x: torch.Tensor) -> torch.Tensor:
    """A ``[N, K]`` view of a zero-padded ``[N, roundup8(K)]`` copy of ``x``: rows start 32-byte
    aligned, so the kernels gather them in place with 16/32-byte loads instead of re-padding on every
    call.  The B200 counterpart of the reference's ``pad_features``
    (/root/reference/tests/cpu/dataset_loader.py:145-160), except that the model still sees K columns."""
    return torch.ops.isplib._b200_pad_features(x)


def isplib_autotune(fn):
    """patch -> call -> unpatch (isplib/__init__.py:204-210); unlike the reference the
    unpatch also happens when ``fn`` raises."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        iSpLibPlugin.patch_pyg()
        try:
            return fn(*args, **kwargs)
        finally:
            iSpLibPlugin.unpatch_pyg()
    return wrapper


__all__ = ["iSpLibPlugin", "isplib_autotune", "SparseTensor", "matmul", "torch", "torch_sparse",
           "__version__"]   # the reference's surface; fused_matmul / pad_features are reached as isplib_b200.*

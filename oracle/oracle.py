"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checker for the iSpLib FusedMM SpMM hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; nothing under ``isplib_b200/`` does.

PARITY UNPINNED: the reference tree holds no golden vector
for this path and the kernel library it links (OnixHoque/FusedMM_Extended, branch
``spmm_variant``, unpinned; /root/reference/configure:2-7) is not in the tree.
Three restatements live here and must agree with each other:

* ``spmm_loops``  -- pure-Python triple loop, the obviously-correct statement of
  /root/reference/csrc/fusedMM.h:77-99 as driven by csrc/fusedmm.cpp:113-203.
* ``spmm_numpy``  -- vectorised numpy (segment reduce), mid-size graphs.
* ``spmm_c``      -- ``fusedmm_oracle.c`` (C/OpenMP, the reference's own C ABI,
  int64 indices) through ctypes; large graphs and the timed CPU baseline.

and, where the real reference code can run, ``oracle/_ref/_fusedmm_cpu.so`` (the
reference's unmodified ``csrc/fusedmm.cpp`` linked against ``fusedmm_oracle.c``)
pins the wrapper and the autograd backward; ``tests/golden/*.npz`` were produced
from it by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfusedmm_oracle.so")
REF_WRAPPER_PATH = os.path.join(HERE, "_ref", "_fusedmm_cpu.so")
REF_GPU_PROTO_PATH = os.path.join(HERE, "_ref", "libref_gpu_proto.so")

# reduction codes of fusedmm_spmm_fw -- csrc/fusedmm.cpp:147-186
SUM, MAX, MIN, MEAN = 0, 1, 2, 3
REDUCE_CODE = {"sum": SUM, "add": SUM, "max": MAX, "min": MIN, "mean": MEAN}

# imsg values the wrapper builds -- csrc/fusedmm.cpp:170,175,181,184
IMSG = {SUM: 0x11102, MAX: 0x21102, MIN: 0x31102, MEAN: 0x13102}

F32_LOWEST = np.float32(np.finfo(np.float32).min)   # numeric_limits<float>::lowest()
F32_MAX = np.float32(np.finfo(np.float32).max)      # numeric_limits<float>::max()


def build(with_ref: bool = False, quiet: bool = True) -> None:
    """Compile the C restatement (and, if /root/reference is present and asked
    for, oracle/_ref).  Building the checker is not using it."""
    targets = ["all"]
    if with_ref and os.path.isdir("/root/reference/csrc"):
        targets.append("ref")
    subprocess.run(["make", "-C", HERE] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None


def load_c() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        i64, f32, p = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
        lib.fusedMM_csr.restype = ctypes.c_int
        lib.fusedMM_csr.argtypes = [ctypes.c_int32, i64, i64, i64, f32, i64, i64, i64,
                                    p, p, p, p, p, i64, p, i64, f32, p, i64, p]
        lib.oracle_arg_backward.restype = ctypes.c_int
        lib.oracle_arg_backward.argtypes = [i64, i64, i64, p, p, p, p, p, p, p]
        lib.oracle_build_csc.restype = ctypes.c_int
        lib.oracle_build_csc.argtypes = [i64, i64, i64, p, p, p, p, p, p]
        lib.oracle_sddmm.restype = ctypes.c_int
        lib.oracle_sddmm.argtypes = [i64, i64, p, p, p, p, ctypes.c_int, p]
        lib.oracle_num_threads.restype = ctypes.c_int
        lib.oracle_spmm_sum_f64.restype = ctypes.c_int
        lib.oracle_spmm_sum_f64.argtypes = [i64, i64, p, p, p, p, ctypes.c_int, p]
        _lib = lib
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _init_out(M: int, K: int, nnz: int, reduction: int):
    """out / arg_out initialisation of the wrapper -- csrc/fusedmm.cpp:147-152,171,177."""
    if reduction == MAX:
        out = np.full((M, K), F32_LOWEST, dtype=np.float32)
    elif reduction == MIN:
        out = np.full((M, K), F32_MAX, dtype=np.float32)
    else:
        out = np.zeros((M, K), dtype=np.float32)
    arg = np.full((M, K), nnz, dtype=np.int64) if reduction in (MAX, MIN) else None
    return out, arg


def _norm_inputs(rowptr, col, value, mat):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    if value is None:
        # isplib/__init__.py:51-57 -- the plugin substitutes fp32 ones
        value = np.ones(col.shape[0], dtype=np.float32)
    value = np.ascontiguousarray(value, dtype=np.float32)
    assert mat.ndim == 2 and rowptr.ndim == 1 and col.ndim == 1
    assert value.shape == col.shape
    return rowptr, col, value, mat


def spmm_c(rowptr, col, value, mat, reduction: int) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """fusedmm_spmm_fw (csrc/fusedmm.cpp:113-203) over the restated C kernel."""
    rowptr, col, value, mat = _norm_inputs(rowptr, col, value, mat)
    M, (N, K), nnz = rowptr.shape[0] - 1, mat.shape, col.shape[0]
    out, arg = _init_out(M, K, nnz, reduction)
    dummy = np.zeros(1, dtype=np.float32)
    lib = load_c()
    pntre = ctypes.c_void_p(rowptr.ctypes.data + 8)  # pntre = pntrb + 1 (csrc/fusedmm.cpp:198)
    st = lib.fusedMM_csr(IMSG[reduction], M, N, K, 1.0, nnz, M, N, _ptr(value), _ptr(col),
                         _ptr(rowptr), pntre,
                         _ptr(dummy), K, _ptr(mat), K, 0.0, _ptr(out), K, _ptr(arg))
    if st != 0:
        raise RuntimeError(f"fusedMM_csr oracle returned status {st}")
    return out, arg


def spmm_loops(rowptr, col, value, mat, reduction: int):
    """Pure-Python statement of the same thing.  Small inputs only."""
    rowptr, col, value, mat = _norm_inputs(rowptr, col, value, mat)
    M, (N, K), nnz = rowptr.shape[0] - 1, mat.shape, col.shape[0]
    out, arg = _init_out(M, K, nnz, reduction)
    for i in range(M):
        b, e = int(rowptr[i]), int(rowptr[i + 1])
        for j in range(b, e):
            a = value[j]
            yr = mat[col[j]]
            for kk in range(K):
                t = np.float32(a * yr[kk])          # VSC_MUL: one fp32 rounding
                if reduction in (SUM, MEAN):
                    out[i, kk] = np.float32(out[i, kk] + t)      # AOP_ADD
                elif reduction == MAX:
                    if t > out[i, kk]:                           # AOP_MAX, strict
                        out[i, kk] = t
                        arg[i, kk] = j
                else:
                    if t < out[i, kk]:                           # AOP_MIN, strict
                        out[i, kk] = t
                        arg[i, kk] = j
        if reduction == MEAN:
            out[i, :] = out[i, :] / np.float32(max(e - b, 1))
    return out, arg


def spmm_numpy(rowptr, col, value, mat, reduction: int):
    """Vectorised restatement (segment reduce).  sum/mean associate differently
    from the sequential loop (np.add.reduceat is pairwise) -- compare with
    tolerance; max/min/arg are exact."""
    rowptr, col, value, mat = _norm_inputs(rowptr, col, value, mat)
    M, (N, K), nnz = rowptr.shape[0] - 1, mat.shape, col.shape[0]
    out, arg = _init_out(M, K, nnz, reduction)
    if nnz == 0 or M == 0:
        return out, arg
    deg = np.diff(rowptr)
    nz = np.nonzero(deg)[0]
    prod = value[:, None] * mat[col]                      # [nnz, K] fp32
    starts = rowptr[:-1][nz]
    if reduction in (SUM, MEAN):
        out[nz] = np.add.reduceat(prod.astype(np.float64), starts, axis=0).astype(np.float32)
        if reduction == MEAN:
            out /= np.maximum(deg, 1).astype(np.float32)[:, None]
    else:
        # fmax / fmin skip NaNs: a NaN product never wins under strict compare (an all-NaN segment stays
        # NaN here and fails `beats` below, so the row keeps the init value and the sentinel)
        red = np.fmax if reduction == MAX else np.fmin
        best = red.reduceat(prod, starts, axis=0)          # [len(nz), K]
        init = F32_LOWEST if reduction == MAX else F32_MAX
        row_of = np.repeat(np.arange(M), deg)
        seg_of = np.searchsorted(nz, row_of)
        hit = prod == best[seg_of]                          # ties -> first edge wins
        eid = np.where(hit, np.arange(nnz)[:, None], nnz)
        first = np.minimum.reduceat(eid, starts, axis=0)
        # an entry only replaces the init value under STRICT compare
        beats = (best > init) if reduction == MAX else (best < init)
        # the stored value is the WINNING entry's product (its bit pattern: +0.0 and -0.0 tie, first seen stays)
        won = np.take_along_axis(prod, np.minimum(first, nnz - 1), axis=0)
        out[nz] = np.where(beats, won, init)
        arg[nz] = np.where(beats, first, nnz)
    return out, arg


# --------------------------------------------------------------------------- #
# backward contracts
# --------------------------------------------------------------------------- #

def build_csc(rowptr, col, N: int):
    """(colptr, csr2csc, row[csr2csc]) -- the stable by-column ordering torch_sparse
    computes and isplib/__init__.py:69-80 consumes."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    M, nnz = rowptr.shape[0] - 1, col.shape[0]
    colptr = np.zeros(N + 1, dtype=np.int64)
    csr2csc = np.zeros(nnz, dtype=np.int64)
    row_t = np.zeros(nnz, dtype=np.int64)
    cursor = np.zeros(max(N, 1), dtype=np.int64)
    st = load_c().oracle_build_csc(M, N, nnz, _ptr(rowptr), _ptr(col), _ptr(colptr),
                                   _ptr(csr2csc), _ptr(row_t), _ptr(cursor))
    if st != 0:
        raise RuntimeError(f"oracle_build_csc status {st}")
    return colptr, csr2csc, row_t


def spmm_backward_sum(rowptr, col, value, grad_out, N: int, impl=spmm_c):
    """grad_mat of sum -- csrc/fusedmm.cpp:258-293: forward over the CSC view with
    value[csr2csc] (isplib/__init__.py:79-80)."""
    rowptr, col, value, grad_out = _norm_inputs(rowptr, col, value, grad_out)
    colptr, csr2csc, row_t = build_csc(rowptr, col, N)
    return impl(colptr, row_t, value[csr2csc], grad_out, SUM)[0]


def spmm_backward_mean(rowptr, col, value, grad_out, N: int, impl=spmm_c):
    """grad_mat of mean -- csrc/fusedmm.cpp:340-383 with the weights of
    isplib/__init__.py:86-93: value[csr2csc] / max(rowcount[row],1)[csr2csc]."""
    rowptr, col, value, grad_out = _norm_inputs(rowptr, col, value, grad_out)
    colptr, csr2csc, row_t = build_csc(rowptr, col, N)
    deg = np.maximum(np.diff(rowptr), 1).astype(np.float32)
    w = value[csr2csc] / deg[row_t]
    return impl(colptr, row_t, w, grad_out, SUM)[0]


def arg_backward(col, value, mat, arg, grad_out, N: int, need_grad_value: bool = False):
    """max/min backward -- csrc/fusedmm.cpp:410-451 / :477-517."""
    col = np.ascontiguousarray(col, dtype=np.int64)
    arg = np.ascontiguousarray(arg, dtype=np.int64)
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    M, K = grad_out.shape
    nnz = col.shape[0]
    val = None if value is None else np.ascontiguousarray(value, dtype=np.float32)
    grad_mat = np.zeros((N, K), dtype=np.float32)
    grad_value = np.zeros(nnz, dtype=np.float32) if need_grad_value else None
    matc = None if mat is None else np.ascontiguousarray(mat, dtype=np.float32)
    st = load_c().oracle_arg_backward(M, K, nnz, _ptr(col), _ptr(val), _ptr(matc), _ptr(arg),
                                      _ptr(grad_out), _ptr(grad_mat), _ptr(grad_value))
    if st != 0:
        raise RuntimeError(f"oracle_arg_backward status {st}")
    return grad_mat, grad_value


def spmm_sum_f64(rowptr, col, value, mat, mean: bool = False) -> np.ndarray:
    """sum / mean with a float64 accumulator (rounded to fp32 once): the yardstick both fp32
    summation orders -- the reference's sequential one and the GPU kernel's -- are measured against."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    val = None if value is None else np.ascontiguousarray(value, dtype=np.float32)
    M, K = rowptr.shape[0] - 1, mat.shape[1]
    out = np.zeros((M, K), dtype=np.float32)
    st = load_c().oracle_spmm_sum_f64(M, K, _ptr(val), _ptr(col), _ptr(rowptr), _ptr(mat), 1 if mean else 0, _ptr(out))
    if st != 0:
        raise RuntimeError(f"oracle_spmm_sum_f64 status {st}")
    return out


def num_threads() -> int:
    return int(load_c().oracle_num_threads())


def sddmm(rowptr, col, a, x, mean_scale: bool = False) -> np.ndarray:
    """grad_value of sum/mean: <a[row(e)], x[col[e]]> (/ max(deg,1)), float64 accumulation."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    a = np.ascontiguousarray(a, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(col.shape[0], dtype=np.float32)
    st = load_c().oracle_sddmm(rowptr.shape[0] - 1, a.shape[1], _ptr(rowptr), _ptr(col), _ptr(a), _ptr(x),
                               1 if mean_scale else 0, _ptr(out))
    if st != 0:
        raise RuntimeError(f"oracle_sddmm status {st}")
    return out


def apply_epilogue(out, bias=None, addend=None, addend_scale: float = 1.0, relu: bool = False) -> np.ndarray:
    """What the reference's callers do to the SpMM result in separate passes, restated:
    GCNConv's `out + bias` and F.relu (tests/cpu/gcn-sparse.py:61-68), GINConv's
    `(1 + eps) * x_i + aggr` (gin-sparse.py:73-78; addend = x, addend_scale = 1 + eps).
    fp32, one rounding per step, in the order  relu((out + scale * addend) + bias)."""
    o = np.asarray(out, dtype=np.float32).copy()
    if addend is not None:
        a = np.asarray(addend, dtype=np.float32)[: o.shape[0]]
        # the kernel uses one fused multiply-add for `scale * addend + out`
        o = (o.astype(np.float64) + np.float64(np.float32(addend_scale)) * a.astype(np.float64)).astype(np.float32)
    if bias is not None:
        o = o + np.asarray(bias, dtype=np.float32)[None, :]
    if relu:
        o = np.maximum(o, np.float32(0))
    return o

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("callers_"))
# what the reference's callers compute around the SpMM (tests/golden/make_golden_callers.py)
CALLER_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("callers_"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device: on a machine without one they are skipped, not failed
    (the product path has no CPU fallback to run them on)."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run on the GPU box with -m gpu")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["value"] = d.get("value", None)
    d["N"] = int(d["N"])
    return d


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.load_c()
    return o


def random_csr(rng, M, N, max_deg, empty_prob=0.0, with_value=True, long_rows=()):
    """Small random CSR (numpy, int64) with sorted columns; duplicates allowed."""
    deg = rng.integers(1, max_deg + 1, size=M)
    if empty_prob > 0:
        deg[rng.random(M) < empty_prob] = 0
    for r, d in long_rows:
        deg[r] = d
    rowptr = np.zeros(M + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    row = np.repeat(np.arange(M), deg)
    col = rng.integers(0, N, size=nnz)
    order = np.lexsort((col, row))
    col = col[order].astype(np.int64)
    val = (rng.random(nnz) * 2 - 1).astype(np.float32) if with_value else None
    return rowptr, col, val


RTOL, ATOL = 1e-5, 1e-6   # north_star: sum/mean within rtol 1e-5 / atol 1e-6 in fp32


def abs_product_sum(rowptr, col, val, mat, mean=False):
    """sum_e |a_e| * |x[col_e, :]| per row in float64: the magnitude the fp32 rounding of a
    reordered sum scales with (the 'condition' of each output element)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    M, K = rowptr.shape[0] - 1, mat.shape[1]
    a = np.ones(col.shape[0]) if val is None else np.abs(np.asarray(val, dtype=np.float64))
    out = np.zeros((M, K))
    deg = np.diff(rowptr)
    nz = np.nonzero(deg)[0]
    if col.shape[0]:
        prod = a[:, None] * np.abs(mat.astype(np.float64))[col]
        out[nz] = np.add.reduceat(prod, rowptr[:-1][nz], axis=0)
    if mean:
        out /= np.maximum(deg, 1)[:, None]
    return out


def assert_sum_close(actual, desired, cond=None):
    """|actual - desired| <= ATOL + RTOL * |desired| -- the north_star tolerance -- where
    `cond` (abs_product_sum) replaces |desired| for elements that suffer cancellation:
    two fp32 summation orders of n terms legitimately differ by ~n*2^-24*sum|t_i|, which
    no order-changing implementation (SIMD-blocked CPU, warp-split GPU) can beat."""
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    scale = np.abs(desired) if cond is None else np.maximum(np.abs(desired), cond)
    err = np.abs(actual - desired)
    bad = err > ATOL + RTOL * scale
    assert not bad.any(), (f"{int(bad.sum())} / {bad.size} elements out of tolerance; "
                           f"max err {err.max():.3e}, worst ratio {(err / (ATOL + RTOL * scale)).max():.2f}")


def grad_cond(oracle, rowptr, col, val, grad_out, N, mean=False):
    """abs_product_sum for the sum/mean BACKWARD (A^T with the backward's weights)."""
    colptr, csr2csc, row_t = oracle.build_csc(rowptr, col, N)
    w = np.ones(col.shape[0], np.float32) if val is None else np.asarray(val, np.float32)
    w = w[csr2csc]
    if mean:
        w = w / np.maximum(np.diff(rowptr), 1)[row_t]
    return abs_product_sum(colptr, row_t, w, grad_out)

"""Property-based tests (hypothesis): random tiny CSR structures -- empty rows, duplicate
entries, rectangular shapes, K from 1 to 70, with and without values, signed zeros and exact
ties -- (a) the three oracle restatements agree (CPU), (b) the CUDA kernels through the C ABI
agree with the oracle for every reduction and a random variant / segment length (GPU)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from conftest import abs_product_sum, assert_sum_close


@st.composite
def csr_problem(draw):
    M = draw(st.integers(1, 24))
    N = draw(st.integers(1, 20))
    K = draw(st.integers(1, 70))
    seed = draw(st.integers(0, 2**31 - 1))
    with_value = draw(st.booleans())
    integer_values = draw(st.booleans())      # exact ties / signed zeros when True
    max_deg = draw(st.sampled_from([0, 1, 3, 40, 300]))
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, max_deg + 1, size=M)
    rowptr = np.zeros(M + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(deg)
    nnz = int(rowptr[-1])
    row = np.repeat(np.arange(M), deg)
    col = rng.integers(0, N, size=nnz)
    order = np.lexsort((col, row))
    col = col[order].astype(np.int64)
    if integer_values:
        val = rng.integers(-2, 3, size=nnz).astype(np.float32) if with_value else None
        mat = rng.integers(-2, 3, size=(N, K)).astype(np.float32)
        mat[mat == 0] = rng.choice([0.0, -0.0], size=int((mat == 0).sum())).astype(np.float32)
    else:
        val = (rng.random(nnz) * 2 - 1).astype(np.float32) if with_value else None
        mat = rng.standard_normal((N, K)).astype(np.float32)
    return rowptr, col, val, mat


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(prob=csr_problem())
def test_oracle_restatements_agree(oracle, prob):
    rowptr, col, val, mat = prob
    if col.shape[0] * mat.shape[1] > 4000:     # keep the pure-Python loops fast
        rowptr, col, val = rowptr[:2], col[: rowptr[1]], (None if val is None else val[: rowptr[1]])
    for code in (oracle.SUM, oracle.MEAN, oracle.MAX, oracle.MIN):
        a, aa = oracle.spmm_loops(rowptr, col, val, mat, code)
        b, ba = oracle.spmm_c(rowptr, col, val, mat, code)
        c, ca = oracle.spmm_numpy(rowptr, col, val, mat, code)
        if code in (oracle.MAX, oracle.MIN):
            assert np.array_equal(a, b) and np.array_equal(aa, ba)
            assert np.array_equal(a, c) and np.array_equal(aa, ca)
        else:
            cond = abs_product_sum(rowptr, col, val, mat, mean=(code == oracle.MEAN))
            assert_sum_close(b, a, cond)
            assert_sum_close(c, a, cond)


@pytest.mark.gpu
@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(prob=csr_problem(), variant_pick=st.integers(0, 10_000), seg_len=st.sampled_from([0, 32, 64, 256]),
       pad=st.booleans())
def test_cuda_kernels_agree_with_oracle(oracle, prob, variant_pick, seg_len, pad):
    import torch
    from isplib_b200 import capi
    rowptr, col, val, mat = prob
    dev = "cuda:0"
    N, K = mat.shape
    rp = torch.from_numpy(rowptr).to(dev).to(torch.int32)
    co = torch.from_numpy(col).to(dev).to(torch.int32)
    va = None if val is None else torch.from_numpy(val).to(dev)
    if pad:                                     # padded rows -> the 16-byte gather path for any K
        Kp = (K + 3) // 4 * 4
        xb = torch.full((N, Kp), float("nan"), device=dev)
        xb[:, :K] = torch.from_numpy(mat).to(dev)
        x = xb[:, :K]
    else:
        x = torch.from_numpy(mat).to(dev)
    plan = capi.Plan(rp, co.numel(), seg_len)
    L = capi.lib()
    for reduce in ("sum", "mean", "max", "min"):
        code = capi.REDUCE_CODE[reduce]
        ok = [v for v in range(L.isplib_b200_variant_count())
              if L.isplib_b200_variant_supported(v, code, K, x.stride(0), K, x.data_ptr(), x.data_ptr())]
        variant = ok[variant_pick % len(ok)]
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, variant)
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, code)
        if reduce in ("max", "min"):
            assert np.array_equal(out.cpu().numpy(), ref), (reduce, capi.variant_names()[variant])
            assert np.array_equal(arg.cpu().numpy(), ref_arg), (reduce, capi.variant_names()[variant])
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean")))

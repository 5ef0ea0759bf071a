"""GPU tests (-m gpu, ONE GPU) of the fused all-gather + SpMM kernel and its grouped plan.

The multi-GPU forward moves the slices of X over NVLink inside the SpMM kernel (the copy CTAs push
the rank's slice into the peers' buffers; isplib_b200_spmm_csr_gather).  Everything but the NVLink hop is exercised here on one device: the
"peers" are ordinary local buffers (RowPartitionedSpMM(emulate=...)), so the copy CTAs, the arrival
flags, the group-ordered work items and the in-kernel merge of split rows all run for real, and the
assembled result must equal the single-GPU oracle: max/min/arg bit-exact, sum/mean in tolerance.
tests/test_gpu_dist.py runs the same operator across real GPUs.
"""
import numpy as np
import pytest
import torch

from conftest import abs_product_sum, assert_sum_close, random_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def capi():
    from isplib_b200 import capi as c
    c.lib()
    return c


def _graph(seed, M, N, max_deg, long_rows=()):
    rng = np.random.default_rng(seed)
    rowptr, col, val = random_csr(rng, M, N, max_deg, empty_prob=0.04, long_rows=long_rows)
    return rng, rowptr, col, val


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
def test_grouped_plan_with_the_plain_kernels(capi, oracle, reduce):
    """A grouped plan only reorders the work items and cuts segments at run boundaries: every
    forward entry point must give the oracle's answer with it."""
    rng, rowptr, col, val = _graph(1, 700, 640, 90, long_rows=[(3, 2500), (11, 640)])
    K = 64
    mat = rng.standard_normal((640, K)).astype(np.float32)
    rp = torch.from_numpy(rowptr).to(DEV).int()
    co = torch.from_numpy(col).to(DEV).int()
    va = torch.from_numpy(val).to(DEV)
    x = torch.from_numpy(mat).to(DEV)
    plan = capi.GroupedPlan(rp, co, [0, 100, 300, 500], [1, 0, 2, 1], 3, seg_len=128)
    assert plan.group_item_end[-1] == plan.info.num_items and sorted(plan.group_item_end) == plan.group_item_end
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    L = capi.lib()
    for v in [-1] + [i for i, nm in enumerate(capi.variant_names()) if nm.startswith(("lean", "seg/w4/u4"))]:
        if v >= 0 and not L.isplib_b200_variant_supported(v, capi.REDUCE_CODE[reduce], K, K, K, x.data_ptr(), x.data_ptr()):
            continue
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, v)
        if reduce in ("max", "min"):
            assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg), f"variant {v}"
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean")))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
@pytest.mark.parametrize("K,groups", [(32, "auto"), (47, "auto"), (128, "auto"), (128, "owners"), (256, "auto")])
def test_fused_gather_emulated_ranks_match_oracle(capi, oracle, world, reduce, K, groups, monkeypatch):
    """K = 128 / 256 with "auto": arrival groups = the 64-wide K tiles (rows whole, plain plan);
    narrower K or "owners": arrival groups = column owners (grouped plan, rows split per group).
    Every emulated rank pushes its slice into the others' buffers with the kernel's copy CTAs, then
    every rank multiplies, waiting on the arrival counters the pushes bumped."""
    from isplib_b200.dist import RowPartitionedSpMM, emulated_step
    monkeypatch.setenv("ISPLIB_B200_DIST_COPY_CTAS", "8")
    monkeypatch.setenv("ISPLIB_B200_DIST_GATHER", groups)
    M = N = 1500 + world          # not a multiple of world: padded slices
    rng, rowptr, col, val = _graph(10 + world, M, N, 60, long_rows=[(5, 1400), (M - 2, 700)])
    rp_t, co_t, va_t = torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(val).to(DEV)
    shared = {}
    ops = [RowPartitionedSpMM(rp_t, co_t, va_t, N, device=DEV, mode="fused", emulate=(world, r, shared),
                              balance="nnz" if world == 4 else "rows") for r in range(world)]
    code = oracle.REDUCE_CODE[reduce]
    for step in range(3):                      # both buffer parities, growing epochs
        mat = rng.standard_normal((N, K)).astype(np.float32)
        x = torch.from_numpy(mat).to(DEV)
        xs = [op.pad_x(x[op.col_range()[0]:op.col_range()[1]]) for op in ops]
        res = emulated_step(ops, xs, reduce)
        out = torch.empty(M, K, device=DEV)
        arg = torch.empty(M, K, dtype=torch.int64, device=DEV)
        for op, (o, a) in zip(ops, res):
            r0, r1 = op.row_range()
            out[r0:r1] = o[: r1 - r0]
            if a is not None:
                arg[r0:r1] = a[: r1 - r0]
        for op in ops:
            op.check_status()
            assert op._tiles_used[K] == (groups == "auto" and K >= 128)
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, code)
        if reduce in ("max", "min"):
            assert np.array_equal(out.cpu().numpy(), ref), f"step {step}"
            assert np.array_equal(arg.cpu().numpy(), ref_arg), f"step {step}"
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean")))


def test_fused_gather_transposed_operator_emulated(capi, oracle, monkeypatch):
    """The sum / mean backward is the same fused kernel on the row partition of A^T (DistSpMM.bwd_op)."""
    from isplib_b200.dist import DistSpMM, emulated_step
    monkeypatch.setenv("ISPLIB_B200_DIST_COPY_CTAS", "4")
    world, K = 4, 64
    M = N = 1203
    rng, rowptr, col, val = _graph(77, M, N, 40, long_rows=[(1, 900)])
    rp_t, co_t, va_t = torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(val).to(DEV)
    shared = {}
    ops = [DistSpMM(rp_t, co_t, va_t, N, device=DEV, mode="fused", emulate=(world, r, shared)) for r in range(world)]
    go = rng.standard_normal((M, K)).astype(np.float32)
    go_d = torch.from_numpy(go).to(DEV)
    for mean in (False, True):
        ts = [op.bwd_op(mean) for op in ops]
        gs = [t.pad_x(go_d[op.fwd.row_range()[0]:op.fwd.row_range()[1]]) for t, op in zip(ts, ops)]
        res = emulated_step(ts, gs, "sum")
        grads = [gx[: op.fwd.col_range()[1] - op.fwd.col_range()[0]] for (gx, _), op in zip(res, ops)]
        bw = oracle.spmm_backward_mean if mean else oracle.spmm_backward_sum
        want = bw(rowptr, col, val, go, N)
        np.testing.assert_allclose(torch.cat(grads).cpu().numpy(), want, rtol=1e-4, atol=1e-4)
        for t in ts:
            t.check_status()

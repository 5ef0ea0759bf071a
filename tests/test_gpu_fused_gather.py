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


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("reduce,K,groups", [("sum", 128, "auto"), ("sum", 47, "auto"), ("mean", 64, "owners"), ("max", 128, "auto")])
def test_fused_gather_epilogue_emulated_ranks(capi, oracle, world, reduce, K, groups, monkeypatch):
    """relu(A x + scale * addend + bias) in the fused gather kernel's final store (tile mode, owner mode with
    rows split per group, ragged width): bit-equal to the epilogue applied to the kernel's own plain result."""
    from isplib_b200.dist import RowPartitionedSpMM, emulated_step, make_epilogue
    monkeypatch.setenv("ISPLIB_B200_DIST_COPY_CTAS", "8")
    monkeypatch.setenv("ISPLIB_B200_DIST_GATHER", groups)
    M = N = 1300 + world
    rng, rowptr, col, val = _graph(40 + world, M, N, 50, long_rows=[(5, 1200)])
    rp_t, co_t, va_t = torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(val).to(DEV)
    shared = {}
    ops = [RowPartitionedSpMM(rp_t, co_t, va_t, N, device=DEV, mode="fused", emulate=(world, r, shared)) for r in range(world)]
    mat = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32)
    x = torch.from_numpy(mat).to(DEV)
    b = torch.from_numpy(bias).to(DEV)
    xs = [op.pad_x(x[op.col_range()[0]:op.col_range()[1]]) for op in ops]
    plain = emulated_step(ops, xs, reduce)
    epis = [make_epilogue(bias=b, addend=op.pad_out(xs[i]), addend_scale=1.5, relu=True) for i, op in enumerate(ops)]
    fused = emulated_step(ops, xs, reduce, epilogues=epis)
    ref_plain, _ = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    for op, (o0, a0), (o1, a1) in zip(ops, plain, fused):
        op.check_status()
        r0, r1 = op.row_range()
        n = r1 - r0
        want = oracle.apply_epilogue(o0.cpu().numpy()[:n], bias=bias, addend=mat[r0:r1], addend_scale=1.5, relu=True)
        got = o1.cpu().numpy()[:n]
        # same reduction result underneath; the epilogue itself adds at most two fp32 roundings
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-5, err_msg=f"rank {op.rank}")
        if a0 is not None:
            assert torch.equal(a0, a1)
        # and against the oracle end to end
        ref = oracle.apply_epilogue(ref_plain[r0:r1], bias=bias, addend=mat[r0:r1], addend_scale=1.5, relu=True)
        np.testing.assert_allclose(got, ref, rtol=2e-6 if reduce == "max" else 1e-4, atol=2e-5 if reduce == "max" else 1e-4)


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
def test_block_kernel_epilogue_after_accumulate(capi, oracle, reduce):
    """The NCCL path's last launch: the remote block ACCUMULATEs into what the local block stored, divides
    (mean) and only then applies bias / addend / ReLU -- through the same block launcher dist.py uses."""
    from isplib_b200.dist import FLAG_ACCUMULATE, RowPartitionedSpMM, make_epilogue
    M = N = 900
    rng, rowptr, col, val = _graph(5, M, N, 40, long_rows=[(7, 800)])
    K = 64
    mat = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32)
    add = rng.standard_normal((M // 2, K)).astype(np.float32)
    shared = {}
    # rank 0 of an emulated world of 2 in NCCL mode: only its two column blocks are used here
    op = RowPartitionedSpMM(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(val).to(DEV),
                            N, device=DEV, mode="nccl", emulate=(2, 0, shared), balance="rows")
    x = torch.from_numpy(mat).to(DEV)
    gathered = torch.cat([op.pad_x(x[c0:c1]) for c0, c1 in (op.col_range(0), op.col_range(1))])
    code = oracle.REDUCE_CODE[reduce]
    inner = 0 if code == 3 else code
    div = op.row_degree if code == 3 else None
    epi = make_epilogue(bias=torch.from_numpy(bias).to(DEV), addend=torch.from_numpy(add).to(DEV), addend_scale=0.5, relu=True)
    outs = []
    for e in ({}, epi):
        out = torch.empty((op.R, K), device=DEV)
        arg = torch.empty((op.R, K), dtype=torch.int64, device=DEV) if inner in (1, 2) else None
        op.block_spmm(inner, op.local, gathered[: op.Rc], out, arg, 0, None, op.nnz, -1)
        op.block_spmm(inner, op.remote, gathered, out, arg, FLAG_ACCUMULATE, div, op.nnz, -1, **e)
        outs.append(out.cpu().numpy())
    r0, r1 = op.row_range()
    want = oracle.apply_epilogue(outs[0][: r1 - r0], bias=bias, addend=add, addend_scale=0.5, relu=True)
    np.testing.assert_allclose(outs[1][: r1 - r0], want, rtol=2e-6, atol=2e-5)
    ref, _ = oracle.spmm_c(rowptr, col, val, mat, code)
    ref = oracle.apply_epilogue(ref[r0:r1], bias=bias, addend=add, addend_scale=0.5, relu=True)
    np.testing.assert_allclose(outs[1][: r1 - r0], ref, rtol=2e-6 if reduce == "max" else 1e-4,
                               atol=2e-5 if reduce == "max" else 1e-4)

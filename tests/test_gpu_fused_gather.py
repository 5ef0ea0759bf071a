"""GPU tests (-m gpu, ONE GPU) of the fused all-gather + SpMM kernel and its grouped plan.

The multi-GPU forward pulls the peers' slices of X over NVLink inside the SpMM kernel
(isplib_b200_spmm_csr_gather).  Everything but the NVLink hop is exercised here on one device: the
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
    narrower K or "owners": arrival groups = column owners (grouped plan, rows split per group)."""
    from isplib_b200.dist import RowPartitionedSpMM
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
        ins = []
        for op in ops:                         # every "rank" publishes its slice first ...
            c0, c1 = op.col_range()
            dst = op.next_input_slice(K)
            dst.zero_()
            dst[: c1 - c0] = x[c0:c1]
            ins.append(dst)
        out = torch.empty(M, K, device=DEV)
        arg = torch.empty(M, K, dtype=torch.int64, device=DEV)
        for op, xin in zip(ops, ins):          # ... then each one runs gather + SpMM in one kernel
            o, a = op.forward(xin, reduce)
            r0, r1 = op.row_range()
            out[r0:r1] = o[: r1 - r0]
            if a is not None:
                arg[r0:r1] = a[: r1 - r0]
        for op in ops:
            op.check_status()
            used_tiles = any(n > 0 for (k_, tm), (_, _, n) in op._gflags.items() if tm)
            assert used_tiles == (groups == "auto" and K >= 128)
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, code)
        if reduce in ("max", "min"):
            assert np.array_equal(out.cpu().numpy(), ref), f"step {step}"
            assert np.array_equal(arg.cpu().numpy(), ref_arg), f"step {step}"
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean")))


def test_fused_gather_autograd_emulated(capi, oracle, monkeypatch):
    """DistSpMM forward + backward (A^T through the same fused kernel) with 4 emulated ranks."""
    from isplib_b200.dist import DistSpMM
    monkeypatch.setenv("ISPLIB_B200_DIST_COPY_CTAS", "4")
    world, K = 4, 64
    M = N = 1203
    rng, rowptr, col, val = _graph(77, M, N, 40, long_rows=[(1, 900)])
    rp_t, co_t, va_t = torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(val).to(DEV)
    shared_f, ops = {}, []
    for r in range(world):
        ops.append(DistSpMM(rp_t, co_t, va_t, N, device=DEV, mode="fused", emulate=(world, r, shared_f)))
    mat = rng.standard_normal((N, K)).astype(np.float32)
    go = rng.standard_normal((M, K)).astype(np.float32)
    x = torch.from_numpy(mat).to(DEV)
    for reduce in ("sum", "mean"):
        # forward: publish all slices, then run every rank
        xs, outs = [], []
        for op in ops:
            c0, c1 = op.fwd.col_range()
            xin = op.fwd.next_input_slice(K)
            xin.zero_()
            xin[: c1 - c0] = x[c0:c1]
            xs.append(xin.clone().requires_grad_(True))
        # the autograd Function copies its (cloned) input into the published slot itself
        for op, xi in zip(ops, xs):
            outs.append(op(xi, reduce))
        ref = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])[0]
        got = torch.cat([o[: op.fwd.row_range()[1] - op.fwd.row_range()[0]] for o, op in zip(outs, ops)]).detach().cpu().numpy()
        assert_sum_close(got, ref, abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean")))
        # backward: the transposed operators also need every rank's grad slice published first
        gts = [op.bwd_op(reduce == "mean") for op in ops]
        for t, op in zip(gts, ops):
            r0, r1 = op.fwd.row_range()
            slot = t.next_input_slice(K)
            slot.zero_()
            slot[: r1 - r0] = torch.from_numpy(go[r0:r1]).to(DEV)
        grads = []
        for t, op in zip(gts, ops):
            gx, _ = t.forward(t.next_input_slice(K), "sum")
            c0, c1 = op.fwd.col_range()
            grads.append(gx[: c1 - c0])
        bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
        want = bw(rowptr, col, val, go, N)
        np.testing.assert_allclose(torch.cat(grads).cpu().numpy(), want, rtol=1e-4, atol=1e-4)

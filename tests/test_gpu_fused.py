"""GPU tests (-m gpu) of what round 2 added around the hot kernel, all through the C ABI or the
registered ops and against the oracle:

* fused caller epilogues (bias, ReLU, scaled addend) in the final store  -- SURVEY 8f rank 1
* auxiliary max/min outputs col[arg] / val[arg] and the streamed backward scatter
* the value-free max/min specialisation and the step-id arg tracking of the lean kernels
* zero-copy padded operands, and the op-layer cache fixes
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


@pytest.fixture(scope="module")
def isplib():
    import isplib as m
    return m


def to_dev(rowptr, col, val, mat):
    rp = torch.from_numpy(np.ascontiguousarray(rowptr)).to(DEV).to(torch.int32)
    co = torch.from_numpy(np.ascontiguousarray(col)).to(DEV).to(torch.int32)
    va = None if val is None else torch.from_numpy(np.ascontiguousarray(val)).to(DEV)
    x = torch.from_numpy(np.ascontiguousarray(mat)).to(DEV)
    return rp, co, va, x


def lean_and_default_variants(capi, reduce, K, x):
    L = capi.lib()
    names = capi.variant_names()
    vs = [-1]
    for v, nm in enumerate(names):
        if nm.startswith("bulk"):
            continue
        if L.isplib_b200_variant_supported(v, capi.REDUCE_CODE[reduce], K, x.stride(0), K, x.data_ptr(), x.data_ptr()):
            vs.append(v)
    return vs


# ------------------------------------------------------------------ fused epilogue, C ABI level
@pytest.mark.parametrize("K", [8, 32, 47, 100, 128, 256])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
@pytest.mark.parametrize("combo", ["bias", "bias+relu", "addend", "all"])
def test_fused_epilogue_matches_oracle(capi, oracle, K, reduce, combo):
    rng = np.random.default_rng(77 + K)
    M = N = 260
    rowptr, col, val = random_csr(rng, M, N, 60, empty_prob=0.05, long_rows=[(3, 1400)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32) if combo in ("bias", "bias+relu", "all") else None
    addend = rng.standard_normal((M, K)).astype(np.float32) if combo in ("addend", "all") else None
    relu = combo in ("bias+relu", "all")
    scale = 1.25
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = capi.Plan(rp, co.numel())
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    want = oracle.apply_epilogue(ref, bias, addend, scale, relu)
    for v in lean_and_default_variants(capi, reduce, K, x):
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, v,
                                 bias=None if bias is None else torch.from_numpy(bias).to(DEV),
                                 addend=None if addend is None else torch.from_numpy(addend).to(DEV),
                                 addend_scale=scale, relu=relu)
        o = out.cpu().numpy()
        if reduce == "max":
            # the reduction itself is exact; the epilogue adds at most two fp32 roundings
            np.testing.assert_allclose(o, want, rtol=2e-6, atol=1e-30, err_msg=f"variant {v}")
            assert np.array_equal(arg.cpu().numpy(), ref_arg)
        else:
            cond = abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean"))
            if relu:   # relu is 1-Lipschitz: the pre-activation bound carries over
                cond = cond + 1.0
            assert_sum_close(o, want, cond + np.abs(want))


# --------------------------------------------------------- auxiliary outputs + streamed backward
@pytest.mark.parametrize("K", [4, 32, 47, 64, 200])
@pytest.mark.parametrize("reduce", ["max", "min"])
@pytest.mark.parametrize("with_value", [True, False])
def test_arg_aux_outputs_and_streamed_backward(capi, oracle, K, reduce, with_value, monkeypatch):
    rng = np.random.default_rng(5 + K)
    M, N = 310, 280
    rowptr, col, val = random_csr(rng, M, N, 50, empty_prob=0.08, with_value=with_value, long_rows=[(9, 900)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    go = rng.standard_normal((M, K)).astype(np.float32)
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = capi.Plan(rp, co.numel())
    nnz = co.numel()
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    rgx, _ = oracle.arg_backward(col, val, mat, ref_arg, go, N, False)
    for v in lean_and_default_variants(capi, reduce, K, x):
        arg_col = torch.full((M, K), -7, dtype=torch.int32, device=DEV)
        arg_val = torch.full((M, K), -7.0, device=DEV) if with_value else None
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, v, arg_col=arg_col, arg_val=arg_val)
        assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg), f"variant {v}"
        a = ref_arg
        has = a != nnz
        want_col = np.where(has, col[np.minimum(a, nnz - 1)], -1)
        assert np.array_equal(arg_col.cpu().numpy(), want_col), f"variant {v}"
        if with_value:
            want_val = np.where(has, val[np.minimum(a, nnz - 1)], np.float32(1))
            assert np.array_equal(arg_val.cpu().numpy(), want_val)
        gx = capi.spmm_arg_backward_aux(arg_col, arg_val, torch.from_numpy(go).to(DEV), N)
        # float atomics add in arrival order: tolerance, not bit-exact
        np.testing.assert_allclose(gx.cpu().numpy(), rgx, rtol=1e-5, atol=1e-5)
        if v == -1:
            # the partition-then-apply scatter (what a grad_x far beyond L2 gets), with slabs small
            # enough that this graph spans many bins, and with the default single bin
            for slab in ("4096", None):
                if slab:
                    monkeypatch.setenv("ISPLIB_B200_BIN_BYTES", slab)
                else:
                    monkeypatch.delenv("ISPLIB_B200_BIN_BYTES", raising=False)
                gb = capi.spmm_arg_backward_aux(arg_col, arg_val, torch.from_numpy(go).to(DEV), N, binned=True)
                np.testing.assert_allclose(gb.cpu().numpy(), rgx, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("K", [8, 32, 64, 104, 128, 256])
@pytest.mark.parametrize("reduce", ["max", "min"])
def test_value_free_max_min_every_variant(capi, oracle, K, reduce):
    """val == NULL (SAGE / GIN drop the values): the multiply-free specialisation, and -- with
    duplicated rows of X -- ties that must resolve to the smallest edge id in every variant."""
    rng = np.random.default_rng(300 + K)
    M, N = 400, 90                                     # few distinct columns: many exact ties
    rowptr, col, _ = random_csr(rng, M, N, 120, empty_prob=0.03, with_value=False, long_rows=[(0, 2100), (7, 640)])
    mat = rng.integers(-3, 4, size=(N, K)).astype(np.float32)     # small integers: ties inside every step
    mat[rng.random((N, K)) < 0.1] = -0.0
    rp, co, va, x = to_dev(rowptr, col, None, mat)
    plan = capi.Plan(rp, co.numel())
    ref, ref_arg = oracle.spmm_c(rowptr, col, None, mat, oracle.REDUCE_CODE[reduce])
    for v in lean_and_default_variants(capi, reduce, K, x):
        out, arg = capi.spmm_csr(reduce, rp, co, None, x, plan, v)
        o = out.cpu().numpy()
        assert np.array_equal(o, ref), f"variant {v}"
        assert np.array_equal(np.signbit(o), np.signbit(ref)), f"variant {v}: -0.0 / +0.0 differ"
        assert np.array_equal(arg.cpu().numpy(), ref_arg), f"variant {v}"


@pytest.mark.parametrize("reduce", ["max", "min"])
def test_weighted_ties_and_signed_zero_every_variant(capi, oracle, reduce):
    rng = np.random.default_rng(11)
    M, N, K = 200, 64, 64
    rowptr, col, _ = random_csr(rng, M, N, 200, with_value=False, long_rows=[(1, 1800)])
    val = rng.choice(np.array([-2.0, -1.0, 0.0, -0.0, 1.0, 2.0], dtype=np.float32), size=col.shape[0])
    mat = rng.integers(-2, 3, size=(N, K)).astype(np.float32)
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = capi.Plan(rp, co.numel())
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    for v in lean_and_default_variants(capi, reduce, K, x):
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, v)
        o = out.cpu().numpy()
        assert np.array_equal(o, ref) and np.array_equal(arg.cpu().numpy(), ref_arg), f"variant {v}"
        assert np.array_equal(np.signbit(o), np.signbit(ref)), f"variant {v}: -0.0 / +0.0 differ"


# ------------------------------------------------------------------------ op level: fused + autograd
def _sparse(isplib, rowptr, col, val, M, N):
    import torch_sparse
    return torch_sparse.SparseTensor(rowptr=torch.from_numpy(rowptr).to(DEV), col=torch.from_numpy(col).to(DEV),
                                     value=None if val is None else torch.from_numpy(val).to(DEV),
                                     sparse_sizes=(M, N), is_sorted=True)


@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("K", [32, 47, 100])
@pytest.mark.parametrize("mode", ["gcn", "gin", "addend"])
def test_fused_op_forward_backward_equals_unfused(isplib, reduce, K, mode):
    """relu(A x + b), (1+eps) x + A x, A x + s y: the fused op and its autograd against the same
    function composed from the plain op and torch ops."""
    import isplib_b200
    rng = np.random.default_rng(9)
    M = N = 300
    rowptr, col, val = random_csr(rng, M, N, 40, empty_prob=0.05, with_value=(mode != "gin"))
    adj = _sparse(isplib, rowptr, col, val, M, N)
    rp, co, va = adj.csr()
    x0 = torch.randn(N, K, device=DEV)
    b0 = torch.randn(K, device=DEV)
    y0 = torch.randn(M, K, device=DEV)
    go = torch.randn(M, K, device=DEV)
    eps = 0.3

    def run(fused):
        x = x0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True)
        y = y0.clone().requires_grad_(True)
        v = None if va is None else va.clone().requires_grad_(True)
        if fused:
            if mode == "gcn":
                out = isplib_b200.fused_matmul(adj.set_value(v, layout="csr") if v is not None else adj, x, reduce, bias=b, relu=True)
            elif mode == "gin":
                out = isplib_b200.fused_matmul(adj, x, reduce, addend=x, addend_scale=1 + eps)
            else:
                out = isplib_b200.fused_matmul(adj.set_value(v, layout="csr") if v is not None else adj, x, reduce, addend=y, addend_scale=0.5, bias=b)
        else:
            ops = torch.ops.isplib
            base = (ops.fusedmm_spmm(None, rp, co, v, None, None, x) if reduce == "sum"
                    else ops.fusedmm_spmm_mean(None, rp, co, v, None, None, None, x))
            if mode == "gcn":
                out = torch.relu(base + b)
            elif mode == "gin":
                out = (1 + eps) * x + base
            else:
                out = base + 0.5 * y + b
        out.backward(go)
        grads = [x.grad]
        if mode != "gin":
            grads.append(b.grad)
        if mode == "addend":
            grads.append(y.grad)
        if v is not None:
            grads.append(v.grad)
        return out.detach(), grads

    of, gf = run(True)
    ou, gu = run(False)
    torch.testing.assert_close(of, ou, rtol=1e-5, atol=1e-5)
    for a, b_ in zip(gf, gu):
        assert a is not None and b_ is not None
        torch.testing.assert_close(a, b_, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("reduce", ["max", "min"])
@pytest.mark.parametrize("with_value", [True, False])
def test_op_arg_backward_through_aux_equals_general(isplib, oracle, reduce, with_value, monkeypatch):
    rng = np.random.default_rng(21)
    M, N, K = 250, 230, 48
    rowptr, col, val = random_csr(rng, M, N, 30, empty_prob=0.1, with_value=with_value)
    rp = torch.from_numpy(rowptr).to(DEV)
    co = torch.from_numpy(col).to(DEV)
    va = None if val is None else torch.from_numpy(val).to(DEV)
    mat = rng.standard_normal((N, K)).astype(np.float32)
    go = rng.standard_normal((M, K)).astype(np.float32)
    op = torch.ops.isplib.fusedmm_spmm_max if reduce == "max" else torch.ops.isplib.fusedmm_spmm_min
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    rgx, _ = oracle.arg_backward(col, val, mat, ref_arg, go, N, False)
    grads = []
    for aux in ("1", "0", "binned"):
        # "binned": the partition-then-apply scatter the op layer uses for a grad_mat beyond 256 MB,
        # forced here by dropping that threshold to 0 MB and shrinking the slabs to 2 KB
        monkeypatch.setenv("ISPLIB_B200_ARG_AUX", "0" if aux == "0" else "1")
        if aux == "binned":
            monkeypatch.setenv("ISPLIB_B200_ARG_BINNED_MIN_MB", "0")
            monkeypatch.setenv("ISPLIB_B200_BIN_BYTES", "2048")
        x = torch.from_numpy(mat).to(DEV).requires_grad_(True)
        out, arg = op(rp, co, va, x)
        assert np.array_equal(out.detach().cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg)
        out.backward(torch.from_numpy(go).to(DEV))
        grads.append(x.grad.cpu().numpy())
        np.testing.assert_allclose(grads[-1], rgx, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------- padded operands (zero copy)
@pytest.mark.parametrize("K", [47, 100, 602])
@pytest.mark.parametrize("reduce", ["sum", "max"])
def test_padded_view_operand_is_used_in_place(isplib, oracle, K, reduce):
    import isplib_b200
    rng = np.random.default_rng(K)
    M = N = 200
    rowptr, col, val = random_csr(rng, M, N, 25)
    mat = rng.standard_normal((N, K)).astype(np.float32)
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    rp, co, va = (torch.from_numpy(a).to(DEV) for a in (rowptr, col, val))
    xp = isplib_b200.pad_features(torch.from_numpy(mat).to(DEV))
    assert xp.shape == (N, K) and xp.stride(0) % 8 == 0 and xp.stride(0) >= K and xp.data_ptr() % 32 == 0
    # garbage in the padding must not leak into the result
    base = xp._base if xp._base is not None else xp
    base[:, K:] = float("nan")
    ops = torch.ops.isplib
    if reduce == "sum":
        out = ops.fusedmm_spmm(None, rp, co, va, None, None, xp)
        assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat))
    else:
        out, arg = ops.fusedmm_spmm_max(rp, co, va, xp)
        assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg)
    # a column-sliced view of a wider matrix (row stride > K, offset != 0) also works, by copy or in place
    wide = torch.randn(N, K + 24, device=DEV)
    view = wide[:, 8:8 + K]
    want, _ = oracle.spmm_c(rowptr, col, val, view.cpu().numpy(), oracle.SUM)
    got = ops.fusedmm_spmm(None, rp, co, va, None, None, view)
    assert_sum_close(got.cpu().numpy(), want, abs_product_sum(rowptr, col, val, view.cpu().numpy()))


# ----------------------------------------------------------------------- ADVICE.md round-1 fixes
def test_value_permutation_cache_survives_address_reuse(isplib, oracle):
    """ADVICE high: edge weights rebuilt every step get the SAME address from the caching allocator
    (version 0 again); the permuted-value cache of the backward must not serve the old weights."""
    rng = np.random.default_rng(3)
    M = N = 500
    K = 16
    rowptr, col, _ = random_csr(rng, M, N, 30, with_value=False)
    rp, co = torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV)
    nnz = col.shape[0]
    go = rng.standard_normal((M, K)).astype(np.float32)
    go_d = torch.from_numpy(go).to(DEV)

    def one_step(step):
        val = rng.standard_normal(nnz).astype(np.float32)
        v = torch.empty(nnz, device=DEV)            # a fresh tensor each step, like learned edge weights
        v.copy_(torch.from_numpy(val))
        ptr = v.data_ptr()
        for reduce in ("sum", "mean"):
            x = torch.randn(N, K, device=DEV, requires_grad=True)
            if reduce == "sum":
                out = torch.ops.isplib.fusedmm_spmm(None, rp, co, v, None, None, x)
            else:
                out = torch.ops.isplib.fusedmm_spmm_mean(None, rp, co, v, None, None, None, x)
            out.backward(go_d)
            bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
            want = bw(rowptr, col, val, go, N)
            np.testing.assert_allclose(x.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-5,
                                       err_msg=f"step {step} {reduce}: stale permuted values")
            del out, x
        return ptr          # every local dies here: the block goes back to the caching allocator

    ptrs = [one_step(i) for i in range(6)]
    assert len(set(ptrs)) < len(ptrs), "the allocator never reused the address; the test did not exercise the hazard"


def test_failed_graph_build_is_not_cached(isplib):
    """ADVICE medium: a build that throws must not leave a half-built cache entry behind."""
    rowptr = torch.tensor([0, 1, 2], device=DEV)
    col = torch.tensor([0, 2**31 + 5], device=DEV)          # does not fit int32
    x = torch.ones(4, 8, device=DEV)
    for _ in range(2):
        with pytest.raises(RuntimeError, match="int32"):
            torch.ops.isplib.fusedmm_spmm(None, rowptr, col, None, None, None, x)


def test_accumulate_after_empty_zero_block(capi, oracle):
    """ADVICE low: ACCUMULATE must treat a previous sentinel as 'no candidate' even when that block
    wrote the EMPTY_ZERO placeholder 0 (a row whose true max is negative)."""
    rowptr1 = np.array([0, 0, 2], dtype=np.int64)           # block 1: row 0 empty
    col1 = np.array([0, 1], dtype=np.int64)
    rowptr2 = np.array([0, 2, 3], dtype=np.int64)           # block 2: row 0 has entries
    col2 = np.array([0, 1, 0], dtype=np.int64)
    mat = -np.abs(np.random.default_rng(0).standard_normal((2, 8))).astype(np.float32) - 1.0   # all negative
    x = torch.from_numpy(mat).to(DEV)
    rp1, co1 = torch.from_numpy(rowptr1).to(DEV).int(), torch.from_numpy(col1).to(DEV).int()
    rp2, co2 = torch.from_numpy(rowptr2).to(DEV).int(), torch.from_numpy(col2).to(DEV).int()
    e1 = torch.tensor([0, 1], dtype=torch.int32, device=DEV)
    e2 = torch.tensor([2, 3, 4], dtype=torch.int32, device=DEV)
    out, arg = capi.spmm_csr("max", rp1, co1, None, x, capi.Plan(rp1, 2), flags=capi.FLAG_EMPTY_ZERO,
                             edge_ids=e1, arg_sentinel=5)
    assert float(out[0].abs().max()) == 0.0 and int(arg[0, 0]) == 5
    out, arg = capi.spmm_csr("max", rp2, co2, None, x, capi.Plan(rp2, 3), out=out, arg_out=arg,
                             flags=capi.FLAG_ACCUMULATE | capi.FLAG_EMPTY_ZERO, edge_ids=e2, arg_sentinel=5)
    want0 = np.maximum(mat[0], mat[1])
    assert np.array_equal(out[0].cpu().numpy(), want0), "a negative true max lost against the placeholder 0"
    assert bool((arg[0] != 5).all())


# ------------------------------------------------------------- node reordering
@pytest.mark.parametrize("order", ["degree", "bfs", "rcm", "random"])
def test_reordering_on_device_commutes_with_cuda_spmm(isplib, oracle, order):
    """(P A P^T)(P x) == P (A x) with the permuted adjacency BUILT ON THE DEVICE (COO -> CSR kernel)
    and multiplied by the CUDA kernels, against the oracle on the unpermuted graph; `bfs` is also
    COMPUTED on the device."""
    import torch_sparse
    from isplib import iSpLibPlugin
    from isplib_b200 import reorder, synth
    g = synth.make_graph(3000, 90_000, law="lognormal", param=1.2, values="uniform", seed=4, device=DEV)
    adj = g.sparse_tensor()
    K = 48
    x = torch.randn(g.n, K, device=DEV)
    perm = {"degree": lambda: reorder.degree_order(adj), "bfs": lambda: reorder.bfs_order(adj),
            "rcm": lambda: reorder.reverse_cuthill_mckee(adj),
            "random": lambda: torch.randperm(g.n, device=DEV)}[order]()
    assert perm.is_cuda and sorted(perm.tolist()) == list(range(g.n))
    adj_p = reorder.permute(adj, perm)
    assert adj_p.csr()[1].is_cuda
    rp, co, va = (t.cpu().numpy() for t in adj.csr())
    xh = x.cpu().numpy()
    p = perm.cpu().numpy()
    iSpLibPlugin.patch_pyg()
    try:
        got_sum = torch_sparse.matmul(adj_p, x[perm], "sum").cpu().numpy()
        got_max = torch_sparse.matmul(adj_p, x[perm], "max").cpu().numpy()
    finally:
        iSpLibPlugin.unpatch_pyg()
    ref_sum = oracle.spmm_c(rp, co, va, xh, oracle.SUM)[0]
    ref_max = oracle.spmm_c(rp, co, va, xh, oracle.MAX)[0]
    assert_sum_close(got_sum, ref_sum[p], abs_product_sum(rp, co, va, xh)[p])
    assert np.array_equal(got_max, ref_max[p])       # the max VALUE is order-independent (arg ids are relabelled)


# ------------------------------------------- caller-level golden vectors from the reference's operator layer
def _caller_cases():
    from conftest import CALLER_CASES
    return CALLER_CASES


@pytest.mark.parametrize("name", _caller_cases())
def test_fused_layers_match_reference_caller_goldens(isplib, name):
    """isplib_b200.nn layers with their fused epilogues (bias + ReLU in the SpMM's final store, GIN's self
    term as the kernel's addend, SAGE's accumulate-GEMM) against what the reference's operator layer + the
    callers' torch ops give on CPU (tests/golden/make_golden_callers.py): outputs and every gradient."""
    import os
    import torch_sparse
    from conftest import GOLDEN_DIR
    from isplib import iSpLibPlugin
    from isplib_b200 import nn as gnn
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    M = z["x"].shape[0]
    val = torch.from_numpy(z["value"]).to(DEV) if "value" in z.files else None
    adj = torch_sparse.SparseTensor(rowptr=torch.from_numpy(z["rowptr"]).to(DEV), col=torch.from_numpy(z["col"]).to(DEV),
                                    value=val, sparse_sizes=(M, M), is_sorted=True)
    t = lambda k: torch.from_numpy(z[k]).to(DEV)
    close = lambda a, k: np.testing.assert_allclose(a.detach().cpu().numpy(), z[k], rtol=2e-4, atol=2e-5, err_msg=k)
    iSpLibPlugin.patch_pyg()
    try:
        # GCN layer + ReLU (PyG order)
        conv = gnn.GCNConv(z["W"].shape[1], z["W"].shape[0], order="linear_first", relu=True).to(DEV)
        with torch.no_grad():
            conv.lin.weight.copy_(t("W"))
            conv.bias.copy_(t("b"))
        x = t("x").requires_grad_(True)
        out = conv(x, adj)
        out.backward(t("grad_out"))
        close(out, "gcn_out"); close(x.grad, "gcn_grad_x"); close(conv.lin.weight.grad, "gcn_grad_W"); close(conv.bias.grad, "gcn_grad_b")
        # GIN aggregation input
        gin = gnn.GINConv(torch.nn.Identity(), eps=float(z["eps"]))
        x = t("x").requires_grad_(True)
        out = gin(x, adj)
        out.backward(t("grad_in"))
        close(out, "gin_out"); close(x.grad, "gin_grad_x")
        # SAGE-mean layer
        sage = gnn.SAGEConv(z["W"].shape[1], z["W"].shape[0], aggr="mean").to(DEV)
        with torch.no_grad():
            sage.lin_l.weight.copy_(t("W")); sage.lin_l.bias.copy_(t("b")); sage.lin_r.weight.copy_(t("Wr"))
        x = t("x").requires_grad_(True)
        out = sage(x, adj)
        out.backward(t("grad_out"))
        close(out, "sage_out"); close(x.grad, "sage_grad_x"); close(sage.lin_l.weight.grad, "sage_grad_Wl")
        close(sage.lin_l.bias.grad, "sage_grad_b"); close(sage.lin_r.weight.grad, "sage_grad_Wr")
    finally:
        iSpLibPlugin.unpatch_pyg()

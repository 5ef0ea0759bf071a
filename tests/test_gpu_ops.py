"""GPU tests of the drop-in surface: torch.ops.isplib.* (same names/schemas as the
reference registers, csrc/fusedmm.cpp:565-570), their autograd, and the plugin
(iSpLibPlugin.patch_pyg -> torch_sparse.matmul), plus full-size property checks on the
Reddit-shaped benchmark graph."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, abs_product_sum, assert_sum_close, grad_cond, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def isplib():
    import isplib as m   # the alias package: `from isplib import *` surface
    return m


def dev_graph(g):
    rowptr = torch.from_numpy(g["rowptr"]).to(DEV)
    col = torch.from_numpy(g["col"]).to(DEV)
    val = None if g["value"] is None else torch.from_numpy(g["value"]).to(DEV)
    return rowptr, col, val


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_ops_forward_backward_vs_reference_autograd(isplib, oracle, name):
    g = load_golden(name)
    cond_f = {m: abs_product_sum(g["rowptr"], g["col"], g["value"], g["mat"], mean=m) for m in (False, True)}
    cond_b = {m: grad_cond(oracle, g["rowptr"], g["col"], g["value"], g["grad_out"], g["N"], mean=m) for m in (False, True)}
    rowptr, col, val = dev_graph(g)
    go = torch.from_numpy(g["grad_out"]).to(DEV)
    ops = torch.ops.isplib
    x = torch.from_numpy(g["mat"]).to(DEV).requires_grad_(True)
    o = ops.fusedmm_spmm(None, rowptr, col, val, None, None, x, None, None)
    o.backward(go)
    assert_sum_close(o.detach().cpu().numpy(), g["sum_out"], cond_f[False])
    assert_sum_close(x.grad.cpu().numpy(), g["sum_grad_mat"], cond_b[False])

    x = torch.from_numpy(g["mat"]).to(DEV).requires_grad_(True)
    o = ops.fusedmm_spmm_mean(None, rowptr, col, val, None, None, None, x, None, None)
    o.backward(go)
    assert_sum_close(o.detach().cpu().numpy(), g["mean_out"], cond_f[True])
    assert_sum_close(x.grad.cpu().numpy(), g["mean_grad_mat"], cond_b[True])

    for red, fn in (("max", ops.fusedmm_spmm_max), ("min", ops.fusedmm_spmm_min)):
        x = torch.from_numpy(g["mat"]).to(DEV).requires_grad_(True)
        v = (val if val is not None else torch.ones(col.numel(), device=DEV)).clone().requires_grad_(True)
        o, arg = fn(rowptr, col, v, x)
        assert arg.dtype == torch.int64 and not arg.requires_grad
        o.backward(go)
        assert np.array_equal(o.detach().cpu().numpy(), g[f"{red}_out"])
        assert np.array_equal(arg.cpu().numpy(), g[f"{red}_arg"])
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{red}_grad_mat"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(v.grad.cpu().numpy(), g[f"{red}_grad_value"], rtol=1e-5, atol=1e-6)


def test_readme_standalone_example(isplib):
    """README.md:95-118 verbatim (on the GPU): patch, build the 3x3 SparseTensor, matmul."""
    from isplib import iSpLibPlugin, SparseTensor
    import torch_sparse
    iSpLibPlugin.patch_pyg()
    try:
        adj_t = SparseTensor(
            row=torch.tensor([2, 0, 1, 0, 0], dtype=torch.int64),
            col=torch.tensor([1, 0, 0, 2, 0], dtype=torch.int64),
            value=torch.tensor([3, 3, 4, 2, -2], dtype=torch.float32),
            sparse_sizes=(3, 3)).to(DEV)
        dense = torch.tensor([[1, 0, 2], [4, 0, 0], [0, 3, 0]], dtype=torch.float32, device=DEV)
        assert torch_sparse.matmul(adj_t, dense).tolist() == [[1, 6, 2], [4, 0, 8], [12, 0, 0]]
        assert torch_sparse.matmul(adj_t, dense, "max").tolist() == [[3, 6, 6], [4, 0, 8], [12, 0, 0]]
        assert torch_sparse.matmul(adj_t, dense, "min")[0].tolist() == [-2, 0, -4]
        np.testing.assert_allclose(torch_sparse.matmul(adj_t, dense, "mean")[0].tolist(), [1 / 3, 2, 2 / 3], rtol=1e-6)
    finally:
        iSpLibPlugin.unpatch_pyg()


def test_cora_shape_via_patch_pyg(isplib, oracle):
    """BASELINE.json config 0: SpMM-sum via patch_pyg() on a Cora-shaped graph, K=64."""
    from isplib import iSpLibPlugin
    from isplib_b200 import synth
    import torch_sparse
    g = synth.make_graph("cora", values="gcn", seed=0)
    assert g.m == 2708 and g.nnz == 10556
    x = torch.randn(g.n, 64, generator=torch.Generator().manual_seed(0))
    adj = g.to(DEV).sparse_tensor()
    xd = x.to(DEV).requires_grad_(True)
    iSpLibPlugin.patch_pyg()
    try:
        out = torch_sparse.matmul(adj, xd, "sum")
        out.sum().backward()
    finally:
        iSpLibPlugin.unpatch_pyg()
    ref, _ = oracle.spmm_c(g.rowptr.numpy(), g.col.numpy(), g.value.numpy(), x.numpy(), oracle.SUM)
    assert_sum_close(out.detach().cpu().numpy(), ref, abs_product_sum(g.rowptr.numpy(), g.col.numpy(), g.value.numpy(), x.numpy()))
    gref = oracle.spmm_backward_sum(g.rowptr.numpy(), g.col.numpy(), g.value.numpy(), np.ones((g.m, 64), np.float32), g.n)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), gref, rtol=1e-5, atol=1e-6)
    # unpatched matmul (stock torch ops) agrees too
    np.testing.assert_allclose(torch_sparse.matmul(adj, xd.detach(), "sum").cpu().numpy(), ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
def test_plugin_matches_oracle_with_autotune(isplib, oracle, reduce):
    """Mid-size graph, big enough to trigger the on-device variant selection."""
    from isplib import iSpLibPlugin
    from isplib_b200 import synth
    import torch_sparse
    g = synth.make_graph(3000, 200_000, law="lognormal", param=1.2, values="uniform", seed=1)
    K = 64
    x = torch.randn(g.n, K, generator=torch.Generator().manual_seed(2))
    adj = g.to(DEV).sparse_tensor()
    xd = x.to(DEV).requires_grad_(True)
    go = torch.randn(g.m, K, generator=torch.Generator().manual_seed(3))
    iSpLibPlugin.patch_pyg()
    try:
        out = torch_sparse.matmul(adj, xd, reduce)
        out2 = torch_sparse.matmul(adj, xd, reduce)      # second call: tuned variant from the cache
        out.backward(go.to(DEV))
    finally:
        iSpLibPlugin.unpatch_pyg()
    assert torch.equal(out, out2), "same inputs must give bit-identical outputs run to run"
    rp, co, va = g.rowptr.numpy(), g.col.numpy(), g.value.numpy()
    code = oracle.REDUCE_CODE[reduce]
    ref, ref_arg = oracle.spmm_c(rp, co, va, x.numpy(), code)
    if reduce in ("max", "min"):
        assert np.array_equal(out.detach().cpu().numpy(), ref)
        gref, _ = oracle.arg_backward(co, va, x.numpy(), ref_arg, go.numpy(), g.n)
        np.testing.assert_allclose(xd.grad.cpu().numpy(), gref, rtol=1e-5, atol=1e-5)
    else:
        assert_sum_close(out.detach().cpu().numpy(), ref, abs_product_sum(rp, co, va, x.numpy(), reduce == "mean"))
        bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
        gref = bw(rp, co, va, go.numpy(), g.n)
        rpt, _, rowt = oracle.build_csc(rp, co, g.n)
        w = va[oracle.build_csc(rp, co, g.n)[1]]
        if reduce == "mean":
            w = w / np.maximum(np.diff(rp), 1)[rowt]
        assert_sum_close(xd.grad.cpu().numpy(), gref, abs_product_sum(rpt, rowt, w, go.numpy()))
    assert torch.ops.isplib._b200_tuned_variant(adj.storage.rowptr(), adj.storage.col(), code, K, True, False) >= 0


def test_graph_cache_eviction_and_staleness(isplib):
    ops = torch.ops.isplib
    ops._b200_cache_clear()
    x = torch.ones(4, 8, device=DEV)
    rowptr = torch.tensor([0, 1, 2], device=DEV)
    col = torch.tensor([0, 1], device=DEV)
    a = ops.fusedmm_spmm(None, rowptr, col, None, None, None, x, None, None)
    assert ops._b200_cache_size() == 1 and a[1, 0] == 1
    col[1] = 3                              # in-place edit bumps the version: entry must be rebuilt
    x[3] = 5.0
    b = ops.fusedmm_spmm(None, rowptr, col, None, None, None, x, None, None)
    assert b[1, 0] == 5
    del rowptr, col
    rowptr2 = torch.tensor([0, 2], device=DEV)
    col2 = torch.tensor([1, 2], device=DEV)
    c = ops.fusedmm_spmm(None, rowptr2, col2, None, None, None, x, None, None)
    assert c.shape == (1, 8) and ops._b200_cache_size() == 1   # dead graph evicted


def test_ops_reject_bad_inputs(isplib):
    ops = torch.ops.isplib
    rowptr = torch.tensor([0, 1], device=DEV)
    col = torch.tensor([0], device=DEV)
    with pytest.raises(RuntimeError, match="float32"):
        ops.fusedmm_spmm(None, rowptr, col, None, None, None, torch.ones(1, 4, device=DEV, dtype=torch.float64), None, None)
    with pytest.raises(RuntimeError, match="2-D"):
        ops.fusedmm_spmm(None, rowptr, col, None, None, None, torch.ones(1, 1, 4, device=DEV), None, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fusedmm_spmm(None, rowptr.cpu(), col.cpu(), None, None, None, torch.ones(1, 4), None, None)


# --------------------------------------------------------------------------------------
# full benchmark size: size-independent properties (no CPU oracle at 114.6M entries)
# --------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def reddit():
    from isplib_b200 import synth, capi
    g = synth.make_graph("reddit", values="uniform", seed=0, device=DEV)
    assert g.m == 232_965 and g.nnz == 114_615_892
    rp = capi.narrow_i64_to_i32(g.rowptr)
    co = capi.narrow_i64_to_i32(g.col)
    plan = capi.Plan(rp, g.nnz)
    return g, rp, co, plan, capi


def test_fullsize_properties(reddit):
    g, rp, co, plan, capi = reddit
    K = 128
    gen = torch.Generator(device=DEV).manual_seed(1)
    x1 = torch.randn(g.n, K, device=DEV, generator=gen)
    # (1) known answer: X = ones  =>  sum row i = sum of the row's values (float64 segment sums)
    ones = torch.ones(g.n, K, device=DEV)
    out, _ = capi.spmm_csr("sum", rp, co, g.value, ones, plan)
    cs = torch.zeros(g.nnz + 1, dtype=torch.float64, device=DEV)
    torch.cumsum(g.value.double(), 0, out=cs[1:])
    rowsum = (cs[g.rowptr[1:]] - cs[g.rowptr[:-1]])
    abssum = torch.zeros(g.nnz + 1, dtype=torch.float64, device=DEV)
    torch.cumsum(g.value.double().abs(), 0, out=abssum[1:])
    scale = (abssum[g.rowptr[1:]] - abssum[g.rowptr[:-1]])
    err = (out.double() - rowsum[:, None]).abs().max(dim=1).values
    assert bool((err <= 1e-6 + 1e-5 * scale).all())
    assert bool((out == out[:, :1]).all())          # every feature lane does the same work
    # (2) linearity: A(2 x1) == 2 (A x1) exactly (scaling by 2 is exact in fp32)
    o1, _ = capi.spmm_csr("sum", rp, co, g.value, x1, plan)
    o2, _ = capi.spmm_csr("sum", rp, co, g.value, 2 * x1, plan)
    assert torch.equal(o2, 2 * o1)
    # (3) every variant gives the same max/min/arg bit for bit, and sum within tolerance
    L = capi.lib()
    base_max, base_arg = capi.spmm_csr("max", rp, co, g.value, x1, plan, variant=0)
    for v in range(1, L.isplib_b200_variant_count()):
        if L.isplib_b200_variant_supported(v, capi.MAX, K, K, K, x1.data_ptr(), x1.data_ptr()):
            m, a = capi.spmm_csr("max", rp, co, g.value, x1, plan, variant=v)
            assert torch.equal(m, base_max) and torch.equal(a, base_arg)
    # (4) arg is a witness: out == val[arg] * x[col[arg]] exactly, arg inside its row, and
    #     max >= mean >= min with no value (mean of the same terms)
    row_lo = g.rowptr[:-1, None]
    row_hi = g.rowptr[1:, None]
    assert bool(((base_arg >= row_lo) & (base_arg < row_hi)).all())
    wit = g.value[base_arg] * torch.gather(x1, 0, g.col[base_arg])
    assert torch.equal(wit, base_max)
    mx, _ = capi.spmm_csr("max", rp, co, None, x1, plan)
    mn, _ = capi.spmm_csr("min", rp, co, None, x1, plan)
    me, _ = capi.spmm_csr("mean", rp, co, None, x1, plan)
    assert bool((mx >= me - 1e-4).all()) and bool((me >= mn - 1e-4).all())
    # (5) idempotence / determinism: same call twice is bit-identical
    o1b, _ = capi.spmm_csr("sum", rp, co, g.value, x1, plan)
    assert torch.equal(o1, o1b)


def test_fullsize_backward_adjoint(reddit):
    """<A x, g> == <x, A^T g> at full size: checks the device-built CSC view end to end."""
    g, rp, co, plan, capi = reddit
    K = 32
    gen = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(g.n, K, device=DEV, generator=gen)
    go = torch.randn(g.m, K, device=DEV, generator=gen)
    colptr, row_t, csr2csc = capi.csr_transpose(rp, co, g.n)
    assert bool((colptr[1:] >= colptr[:-1]).all()) and int(colptr[-1]) == g.nnz
    assert bool((torch.sort(csr2csc).values == torch.arange(g.nnz, device=DEV, dtype=torch.int32)).all())
    plan_t = capi.Plan(colptr, g.nnz)
    vt = capi.permute_values(g.value, csr2csc, row_t, rp, False)
    y, _ = capi.spmm_csr("sum", rp, co, g.value, x, plan)
    gx, _ = capi.spmm_csr("sum", colptr, row_t, vt, go, plan_t)
    lhs = float((y.double() * go.double()).sum())
    rhs = float((x.double() * gx.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs), float((y.double().abs() * go.double().abs()).sum()) * 1e-2)


# --------------------------------------------------------------------------------------
# callers: the GNN layers of the reference's benchmark scripts, patched vs stock matmul
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("model_name", ["gcn", "sage-mean", "sage-sum", "gin"])
def test_gnn_training_step_patched_equals_stock(isplib, model_name):
    """One full training step (forward, nll_loss, backward) of the 2-layer models the reference
    benchmarks (tests/cpu/gcn-sparse.py, graphSAGE-sparse.py, gin-sparse.py): with the plugin
    active (CUDA kernels) vs inactive (stock torch-op matmul) -> same loss and gradients."""
    import copy
    import torch.nn.functional as F
    from isplib import iSpLibPlugin
    from isplib_b200 import nn as gnn, synth
    torch.manual_seed(0)
    g = synth.make_graph(1500, 40_000, law="lognormal", param=1.0, values="gcn" if model_name == "gcn" else None, seed=4)
    adj = g.to(DEV).sparse_tensor()
    feat, hidden, classes = 24, 32, 7
    x = torch.randn(g.n, feat, device=DEV)
    y = torch.randint(0, classes, (g.n,), device=DEV)
    if model_name == "gcn":
        model = gnn.GCN(feat, hidden, classes, dropout=0.0)
    elif model_name.startswith("sage"):
        model = gnn.GraphSAGE(feat, hidden, classes, aggr=model_name.split("-")[1], dropout=0.0)
    else:
        model = gnn.GIN(feat, hidden, classes)
    model = model.to(DEV).eval() if model_name == "gin" else model.to(DEV)   # GIN: no dropout randomness
    ref_model = copy.deepcopy(model)

    def step(m):
        out = m(x, adj)
        lp = out if model_name != "gin" else F.log_softmax(out, dim=1)
        loss = F.nll_loss(lp, y)
        loss.backward()
        return loss.item(), [p.grad.clone() for p in m.parameters()]

    iSpLibPlugin.patch_pyg()
    try:
        loss_a, grads_a = step(model)
    finally:
        iSpLibPlugin.unpatch_pyg()
    loss_b, grads_b = step(ref_model)
    assert abs(loss_a - loss_b) <= 1e-4 * max(1.0, abs(loss_b))
    for ga, gb in zip(grads_a, grads_b):
        torch.testing.assert_close(ga, gb, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_value_gradient_of_sum_and_mean(isplib, oracle, reduce):
    """grad w.r.t. the adjacency values (learnable edge weights): undefined in the reference
    (csrc/fusedmm.cpp:268-272, 349-353), provided here as an SDDMM; checked against the
    analytic derivative <grad_out[row(e)], X[col(e)]> (/deg) and against grad_mat's oracle."""
    from isplib_b200 import synth
    g = synth.make_graph(900, 30_000, law="lognormal", param=1.2, values="uniform", seed=8)
    K = 48
    x = torch.randn(g.n, K, generator=torch.Generator().manual_seed(1))
    go = torch.randn(g.m, K, generator=torch.Generator().manual_seed(2))
    rowptr, col = g.rowptr.to(DEV), g.col.to(DEV)
    val = g.value.to(DEV).requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    ops = torch.ops.isplib
    if reduce == "sum":
        out = ops.fusedmm_spmm(None, rowptr, col, val, None, None, xd, None, None)
    else:
        out = ops.fusedmm_spmm_mean(None, rowptr, col, val, None, None, None, xd, None, None)
    out.backward(go.to(DEV))
    ref = oracle.sddmm(g.rowptr.numpy(), g.col.numpy(), go.numpy(), x.numpy(), reduce == "mean")
    np.testing.assert_allclose(val.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
    bw = oracle.spmm_backward_sum if reduce == "sum" else oracle.spmm_backward_mean
    gref = bw(g.rowptr.numpy(), g.col.numpy(), g.value.numpy(), go.numpy(), g.n)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), gref, rtol=1e-4, atol=1e-4)


# --------------------------------------------------------------------------------------
# ingest formats: edge_index -> CSR on the device, Matrix Market round trip
# --------------------------------------------------------------------------------------
def test_from_edge_index_matches_host_sparse_tensor(isplib):
    from isplib import SparseTensor
    from isplib_b200 import io
    gen = torch.Generator().manual_seed(5)
    M, N, E = 300, 200, 5000
    ei = torch.stack([torch.randint(0, M, (E,), generator=gen), torch.randint(0, N, (E,), generator=gen)])
    ei[:, 10] = ei[:, 3]                      # a duplicate entry: input order must be kept (stable)
    w = torch.randn(E, generator=gen)
    ref = SparseTensor(row=ei[0], col=ei[1], value=w, sparse_sizes=(M, N))          # host: stable argsort
    got = io.from_edge_index(ei, w, M, N, device=DEV)
    for a, b in zip(got.csr(), ref.csr()):
        assert torch.equal(a.cpu(), b)
    assert got.sparse_sizes() == (M, N)
    nov = io.from_edge_index(ei, None, M, N, device=DEV)
    assert nov.storage.value() is None and torch.equal(nov.storage.col().cpu(), ref.storage.col())
    # README.md:105-110 fixture through the device builder
    adj = io.from_edge_index(torch.tensor([[2, 0, 1, 0, 0], [1, 0, 0, 2, 0]]), torch.tensor([3., 3., 4., 2., -2.]), 3, 3, DEV)
    assert adj.csr()[0].tolist() == [0, 3, 4, 5] and adj.csr()[2].tolist() == [3, -2, 2, 4, 3]
    with pytest.raises(Exception):
        io.from_edge_index(torch.tensor([[0, 5], [0, 1]]), None, 3, 3, DEV)         # row 5 out of range


def test_matrix_market_round_trip_and_spmm(isplib, oracle, tmp_path):
    from isplib import iSpLibPlugin
    from isplib_b200 import io, synth
    import torch_sparse
    g = synth.make_graph(400, 6000, law="lognormal", param=1.0, values="uniform", seed=6)
    # Matrix Market sums duplicates; make the graph duplicate-free so the round trip is exact
    row = torch.repeat_interleave(torch.arange(g.m), g.rowptr[1:] - g.rowptr[:-1])
    key = torch.unique(row * g.n + g.col)
    row, col = key // g.n, key % g.n
    val = torch.rand(key.numel(), generator=torch.Generator().manual_seed(1)) + 0.5
    from isplib import SparseTensor
    adj = SparseTensor(row=row, col=col, value=val, sparse_sizes=(g.m, g.n))
    path = str(tmp_path / "graph.mtx")
    io.write_mtx(adj, path)
    back = io.read_mtx(path, device=DEV)
    for a, b in zip(back.csr(), adj.csr()):
        if a.dtype.is_floating_point:
            torch.testing.assert_close(a.cpu(), b, rtol=1e-6, atol=0)
        else:
            assert torch.equal(a.cpu(), b)
    x = torch.randn(g.n, 32, generator=torch.Generator().manual_seed(2))
    iSpLibPlugin.patch_pyg()
    try:
        out = torch_sparse.matmul(back, x.to(DEV), "sum")
    finally:
        iSpLibPlugin.unpatch_pyg()
    rp, co, va = [t.numpy() for t in adj.csr()]
    ref, _ = oracle.spmm_c(rp, co, back.storage.value().cpu().numpy(), x.numpy(), oracle.SUM)
    assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rp, co, va, x.numpy()))


def test_empty_row_mode_zero_via_plugin(isplib):
    """PyG's aggr='max' expects 0 for isolated nodes (torch_sparse); the reference wrapper leaves
    lowest(); both are available, the reference behaviour is the default."""
    from isplib import iSpLibPlugin, SparseTensor
    import torch_sparse
    adj = SparseTensor(row=torch.tensor([0, 0, 2]), col=torch.tensor([0, 1, 1]), value=torch.tensor([1., -2., 3.]),
                       sparse_sizes=(4, 2)).to(DEV)
    x = torch.tensor([[1., -1.], [2., 5.]], device=DEV)
    iSpLibPlugin.patch_pyg()
    try:
        ref_mode = torch_sparse.matmul(adj, x, "max")
        iSpLibPlugin.set_empty_row_mode("zero")
        zero_mode = torch_sparse.matmul(adj, x, "max")
        zmin = torch_sparse.matmul(adj, x, "min")
    finally:
        iSpLibPlugin.set_empty_row_mode("reference")
        iSpLibPlugin.unpatch_pyg()
    lowest = torch.finfo(torch.float32).min
    assert ref_mode.tolist() == [[1., -1.], [lowest, lowest], [6., 15.], [lowest, lowest]]
    assert zero_mode.tolist() == [[1., -1.], [0., 0.], [6., 15.], [0., 0.]]
    assert zmin.tolist() == [[-4., -10.], [0., 0.], [6., 15.], [0., 0.]]


def test_ops_are_cuda_graph_capturable(isplib):
    """After the first call per graph (plan build + variant selection synchronise once), the
    op is allocation-pool friendly and sync-free, so a forward+backward can be captured in a
    CUDA graph and replayed -- the way to run launch-bound small graphs (Cora-shape)."""
    from isplib_b200 import synth
    g = synth.make_graph("cora", values="gcn", seed=0).to(DEV)
    rowptr, col, val = g.rowptr, g.col, g.value
    ops = torch.ops.isplib
    x = torch.randn(g.n, 64, device=DEV, requires_grad=True)
    go = torch.randn(g.m, 64, device=DEV)

    def step():
        x.grad = None
        out = ops.fusedmm_spmm(None, rowptr, col, val, None, None, x, None, None)
        mx, _ = ops.fusedmm_spmm_max(rowptr, col, val, x)
        (out + mx).backward(go)
        return out, mx

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):                      # warm-up: plan, CSC view, variant selection
            step()
    torch.cuda.current_stream().wait_stream(s)
    ref_out, ref_mx = [t.detach().clone() for t in step()]
    ref_grad = x.grad.detach().clone()
    graph = torch.cuda.CUDAGraph()
    x.grad = None
    with torch.cuda.graph(graph):
        out, mx = step()
    with torch.no_grad():
        x.copy_(x * 1.0)                        # same values; replay must recompute from x
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref_out) and torch.equal(mx, ref_mx)
    torch.testing.assert_close(x.grad, ref_grad, rtol=1e-5, atol=1e-6)


def test_ops_from_concurrent_threads_and_streams(isplib, oracle):
    """The op layer is called from the Python thread (forward) and from autograd worker
    threads (backward) in real training (SURVEY 8b 'Threading'): hammer it from 4 threads, each
    on its own CUDA stream, two graphs shared between them."""
    import threading
    from isplib_b200 import synth
    graphs = [synth.make_graph(700, 25_000, law="lognormal", param=1.1, values="uniform", seed=s).to(DEV) for s in (1, 2)]
    K = 32
    x = torch.randn(700, K, device=DEV)
    refs = []
    for g in graphs:
        r, _ = oracle.spmm_c(g.rowptr.cpu().numpy(), g.col.cpu().numpy(), g.value.cpu().numpy(), x.cpu().numpy(), oracle.MAX)
        refs.append(torch.from_numpy(r).to(DEV))
    errors = []

    def worker(tid):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for it in range(20):
                    g = graphs[(tid + it) % 2]
                    xr = x.clone().requires_grad_(True)
                    out, _ = torch.ops.isplib.fusedmm_spmm_max(g.rowptr, g.col, g.value, xr)
                    s = torch.ops.isplib.fusedmm_spmm(None, g.rowptr, g.col, g.value, None, None, xr, None, None)
                    (out.sum() + s.sum()).backward()
                    if not torch.equal(out.detach(), refs[(tid + it) % 2]):
                        errors.append((tid, it, "mismatch"))
            stream.synchronize()
        except Exception as ex:   # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors[:3]

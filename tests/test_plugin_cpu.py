"""CPU tests of the Python surface (no kernels run): same names as the reference's
isplib/__init__.py, patch/unpatch stack semantics, loud failure on CPU tensors."""
import sys

import pytest
import torch


def test_surface_names():
    import isplib
    ns = {}
    exec("from isplib import *", ns)
    for name in ("iSpLibPlugin", "isplib_autotune", "SparseTensor", "matmul", "torch", "torch_sparse"):
        assert name in ns, name
    P = isplib.iSpLibPlugin
    for attr in ("backup", "value_cache", "cache", "row_cache", "is_cached", "value_cached"):   # isplib/__init__.py:35-40
        assert hasattr(P, attr)
    for op in ("fusedmm_spmm", "fusedmm_spmm_mean", "fusedmm_spmm_max", "fusedmm_spmm_min", "performDummySpMM"):
        assert hasattr(torch.ops.isplib, op)            # csrc/fusedmm.cpp:565-570
    torch.ops.isplib.performDummySpMM(0)


def test_op_schemas_match_reference_signatures():
    s = torch.ops.isplib.fusedmm_spmm.default._schema
    assert [a.name for a in s.arguments] == ["row", "rowptr", "col", "value", "colptr", "csr2csc", "mat",
                                             "value_index_select", "row_index_select"]   # csrc/fusedmm.cpp:520-527
    s = torch.ops.isplib.fusedmm_spmm_mean.default._schema
    assert [a.name for a in s.arguments] == ["row", "rowptr", "col", "value", "rowcount", "colptr", "csr2csc", "mat",
                                             "new_row", "new_rowcount"]                    # csrc/fusedmm.cpp:535-543
    for op in (torch.ops.isplib.fusedmm_spmm_max, torch.ops.isplib.fusedmm_spmm_min):
        s = op.default._schema
        assert [a.name for a in s.arguments] == ["rowptr", "col", "value", "mat"] and len(s.returns) == 2


def test_patch_unpatch_is_a_stack():
    import isplib
    ts = sys.modules["torch_sparse"]
    P = isplib.iSpLibPlugin
    orig_mm, orig_sparse_mm = ts.matmul, torch.sparse.mm
    P.unpatch_pyg()                                   # no-op when not patched (isplib/__init__.py:190)
    assert ts.matmul is orig_mm
    P.patch_pyg()
    assert ts.matmul == P.spmm and torch.sparse.mm == P.spmm and P.is_patched()
    P.patch_pyg()
    P.unpatch_pyg()
    assert ts.matmul == P.spmm                        # still patched once
    P.unpatch_pyg()
    assert ts.matmul is orig_mm and torch.sparse.mm is orig_sparse_mm and not P.is_patched()


def test_process_group_follows_the_patch_stack():
    """patch_pyg(group=g) scopes nest: an inner patch without a group does not lose the outer group."""
    import isplib
    P = isplib.iSpLibPlugin
    g = object()                                      # only stored and handed on; never used without a partition
    assert P.dist_group is None
    P.patch_pyg(group=g)
    assert P.dist_group is g
    P.patch_pyg()                                     # e.g. an @isplib_autotune function called inside
    assert P.dist_group is None
    P.unpatch_pyg()
    assert P.dist_group is g and P.is_patched()
    P.unpatch_pyg()
    assert P.dist_group is None and not P.is_patched()
    P.unpatch_pyg()                                   # still a no-op when nothing is patched
    assert P.dist_group is None


def test_decorator_unpatches_even_on_exception():
    import isplib
    ts = sys.modules["torch_sparse"]
    orig = ts.matmul

    @isplib.isplib_autotune
    def boom():
        assert ts.matmul == isplib.iSpLibPlugin.spmm
        raise KeyError("x")

    with pytest.raises(KeyError):
        boom()
    assert ts.matmul is orig

    @isplib.isplib_autotune
    def fine(a, b=2):
        return a + b
    assert fine(1, b=3) == 4 and ts.matmul is orig


def test_cpu_tensors_fail_loudly_no_fallback():
    import isplib
    ST = sys.modules["torch_sparse"].SparseTensor
    adj = ST(row=torch.tensor([0, 1]), col=torch.tensor([1, 0]), sparse_sizes=(2, 2))
    x = torch.ones(2, 4)
    isplib.iSpLibPlugin.patch_pyg()
    try:
        with pytest.raises(RuntimeError, match="CUDA"):
            sys.modules["torch_sparse"].matmul(adj, x, "sum")
        with pytest.raises(ValueError):
            isplib.iSpLibPlugin.spmm(adj, x, "median")
        # torch.sparse.mm on genuine torch sparse tensors still works while patched
        sp = torch.eye(2).to_sparse()
        assert torch.equal(torch.sparse.mm(sp, x), x)
    finally:
        isplib.iSpLibPlugin.unpatch_pyg()
    # unpatched: the stock (shim) matmul runs on CPU
    assert sys.modules["torch_sparse"].matmul(adj, x, "sum").tolist() == [[1] * 4, [1] * 4]


def test_shim_sparse_tensor_storage_fields():
    ST = sys.modules["torch_sparse"].SparseTensor
    adj = ST(row=torch.tensor([2, 0, 1, 0, 0]), col=torch.tensor([1, 0, 0, 2, 0]),
             value=torch.tensor([3., 3., 4., 2., -2.]), sparse_sizes=(3, 3))      # README.md:105-110
    rowptr, col, value = adj.csr()
    assert rowptr.tolist() == [0, 3, 4, 5] and col.tolist() == [0, 0, 2, 0, 1]
    assert value.tolist() == [3, -2, 2, 4, 3]                                   # stable: 3 stays before -2
    st = adj.storage
    assert st._rowcount is None and st._csr2csc is None and st._colptr is None   # lazily filled
    assert st.rowcount().tolist() == [3, 1, 1] and st.colptr().tolist() == [0, 3, 4, 5]
    assert st.csr2csc().tolist() == [0, 1, 3, 4, 2]
    assert adj.set_value(None).storage.value() is None and adj.storage.value() is not None
    t = adj.t()
    assert t.sparse_sizes() == (3, 3) and t.nnz() == 5


def test_missing_extension_is_an_import_error(tmp_path, monkeypatch):
    import isplib_b200
    monkeypatch.setattr(isplib_b200, "_PKG_DIR", str(tmp_path))
    with pytest.raises(ImportError, match="no CPU implementation"):
        isplib_b200._load_extension()


def test_synth_generators():
    from isplib_b200 import synth
    g = synth.make_graph("cora", values="gcn", seed=0)
    assert (g.m, g.nnz) == (2708, 10556) and int(g.rowptr[-1]) == g.nnz
    deg = g.rowptr[1:] - g.rowptr[:-1]
    assert int(deg.min()) >= 1 and g.max_degree == int(deg.max())
    row = torch.repeat_interleave(torch.arange(g.m), deg)
    key = row * g.n + g.col
    assert bool((key[1:] >= key[:-1]).all())                      # columns sorted inside rows
    g2 = synth.make_graph("cora", values="gcn", seed=0)
    assert torch.equal(g.col, g2.col) and torch.equal(g.value, g2.value)
    ge = synth.make_graph(1000, 20_000, law="zipf", param=2.1, empty_frac=0.05, values=None, seed=3)
    dege = ge.rowptr[1:] - ge.rowptr[:-1]
    assert int(dege.sum()) == 20_000 and int((dege == 0).sum()) == 50 and ge.value is None
    small = synth.make_graph("reddit", scale=1e-3, seed=0)
    assert small.m == 233 and abs(small.nnz - 114_616) <= 1
    b = synth.algorithmic_bytes(232_965, 114_615_892, 128, True)
    assert abs(b / 1e9 - 59.72) < 0.01


def test_gcn_layer_orders_are_the_same_function():
    """(A x) W == A (x W): the aggregate-first order of isplib_b200.nn.GCNConv is PyG's
    GCNConv up to fp32 reassociation (stock CPU matmul, plugin not active)."""
    import isplib_b200  # noqa: F401
    from isplib_b200 import nn as gnn, synth
    torch.manual_seed(0)
    g = synth.make_graph(300, 4000, values="gcn", seed=2)
    adj = g.sparse_tensor()
    x = torch.randn(g.n, 12)
    a = gnn.GCNConv(12, 20, order="linear_first")
    b = gnn.GCNConv(12, 20, order="aggregate_first")
    b.load_state_dict(a.state_dict())
    assert not a.aggregate_first and b.aggregate_first and gnn.GCNConv(12, 20).aggregate_first
    assert not gnn.GCNConv(20, 12).aggregate_first
    torch.testing.assert_close(a(x, adj), b(x, adj), rtol=1e-4, atol=1e-5)


def test_reordering_commutes_with_spmm():
    """(P A P^T)(P x) == P (A x): isplib_b200.reorder on the stock CPU matmul."""
    import isplib_b200  # noqa: F401
    from isplib_b200 import reorder, synth
    g = synth.make_graph(200, 3000, values="uniform", seed=3)
    adj = g.sparse_tensor()
    x = torch.randn(g.n, 5)
    ref = sys.modules["torch_sparse"].matmul(adj, x, "sum")
    for perm in (reorder.degree_order(adj), reorder.reverse_cuthill_mckee(adj), reorder.bfs_order(adj), torch.randperm(g.n)):
        assert sorted(perm.tolist()) == list(range(g.n))
        adj_p = reorder.permute(adj, perm)
        out_p = sys.modules["torch_sparse"].matmul(adj_p, x[perm], "sum")
        torch.testing.assert_close(out_p, ref[perm], rtol=1e-5, atol=1e-5)
        mx = sys.modules["torch_sparse"].matmul(adj_p, x[perm], "max")
        torch.testing.assert_close(mx, sys.modules["torch_sparse"].matmul(adj, x, "max")[perm])


def test_relabel_by_degree_is_a_symmetric_permutation():
    """synth.relabel_by_degree returns P A P^T with non-increasing row degrees; (P A P^T)(P x) = P (A x)."""
    import numpy as np
    from isplib_b200 import synth
    from oracle import oracle
    g = synth.make_graph(300, 5000, law="lognormal", param=1.2, seed=3, device="cpu", values="uniform")
    rp, co, va = synth.relabel_by_degree(g.rowptr, g.col, g.value, g.n)
    deg = (rp[1:] - rp[:-1]).numpy()
    assert (np.diff(deg) <= 0).all() and int(rp[-1]) == g.nnz
    for i in range(0, 300, 37):                      # columns stay sorted inside a row
        seg = co[int(rp[i]):int(rp[i + 1])].numpy()
        assert (np.diff(seg) >= 0).all()
    old_deg = (g.rowptr[1:] - g.rowptr[:-1])
    perm = torch.argsort(old_deg, descending=True, stable=True).numpy()
    x = np.random.default_rng(0).standard_normal((300, 8)).astype(np.float32)
    ref, _ = oracle.spmm_c(g.rowptr.numpy(), g.col.numpy(), g.value.numpy(), x, oracle.SUM)
    got, _ = oracle.spmm_c(rp.numpy(), co.numpy(), va.numpy(), x[perm], oracle.SUM)
    np.testing.assert_allclose(got, ref[perm], rtol=1e-5, atol=1e-5)

"""CPU tests of the oracle itself: the three restatements agree with each other, with
the committed golden vectors (produced by the reference's own wrapper + autograd, see
tests/golden/make_golden.py) and with the two input fixtures the reference tree holds
(README.md:105-116, gpu/fusedmm.cu:60-118)."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, abs_product_sum, assert_sum_close, load_golden, random_csr

RTOL, ATOL = 1e-5, 1e-6   # north_star tolerance for sum/mean (fp32)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_c_oracle_matches_reference_wrapper_golden(oracle, name):
    g = load_golden(name)
    for red, code in (("sum", oracle.SUM), ("mean", oracle.MEAN), ("max", oracle.MAX), ("min", oracle.MIN)):
        out, arg = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], code)
        if red in ("max", "min"):
            # bit-exact: the golden run used the same C kernel under the reference wrapper
            assert np.array_equal(out, g[f"{red}_out"])
            assert np.array_equal(arg, g[f"{red}_arg"])
        else:
            np.testing.assert_allclose(out, g[f"{red}_out"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_loops_numpy_and_c_agree(oracle, name):
    g = load_golden(name)
    for code in (oracle.SUM, oracle.MEAN, oracle.MAX, oracle.MIN):
        a, aa = oracle.spmm_loops(g["rowptr"], g["col"], g["value"], g["mat"], code)
        b, ba = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], code)
        c, ca = oracle.spmm_numpy(g["rowptr"], g["col"], g["value"], g["mat"], code)
        if code in (oracle.MAX, oracle.MIN):
            assert np.array_equal(a, b) and np.array_equal(aa, ba)
            assert np.array_equal(a, c) and np.array_equal(aa, ca)
        else:
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(a, c, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backward_restatements_match_reference_autograd(oracle, name):
    """oracle.spmm_backward_* / arg_backward vs gradients produced by the reference's
    own autograd Functions (csrc/fusedmm.cpp:258-293, 340-383, 410-451, 477-517)."""
    g = load_golden(name)
    N = g["N"]
    gs = oracle.spmm_backward_sum(g["rowptr"], g["col"], g["value"], g["grad_out"], N)
    np.testing.assert_allclose(gs, g["sum_grad_mat"], rtol=RTOL, atol=ATOL)
    gm = oracle.spmm_backward_mean(g["rowptr"], g["col"], g["value"], g["grad_out"], N)
    np.testing.assert_allclose(gm, g["mean_grad_mat"], rtol=RTOL, atol=ATOL)
    for red in ("max", "min"):
        val = g["value"] if g["value"] is not None else np.ones(g["col"].shape[0], np.float32)
        gx, gv = oracle.arg_backward(g["col"], val, g["mat"], g[f"{red}_arg"], g["grad_out"], N, True)
        np.testing.assert_allclose(gx, g[f"{red}_grad_mat"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(gv, g[f"{red}_grad_value"], rtol=RTOL, atol=ATOL)


def test_readme_fixture_known_answers(oracle):
    """README.md:105-116: 3x3 with a duplicate (0,0) entry (3 and -2)."""
    g = load_golden("readme_3x3")
    s, _ = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.SUM)
    assert s.tolist() == [[1, 6, 2], [4, 0, 8], [12, 0, 0]]
    mx, amx = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.MAX)
    assert mx.tolist() == [[3, 6, 6], [4, 0, 8], [12, 0, 0]]
    assert amx.tolist() == [[0, 2, 0], [3, 3, 3], [4, 4, 4]]
    mn, amn = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.MIN)
    assert mn[0, 0] == -2 and mn[0, 2] == -4
    # (0,1): products are +0.0 (3*0), -0.0 (-2*0), +0.0 (2*0): all compare equal -> first edge wins
    assert mn[0, 1] == 0 and amn[0, 1] == 0 and not np.signbit(mn[0, 1])
    me, _ = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.MEAN)
    np.testing.assert_allclose(me[0], [1 / 3, 2, 2 / 3], rtol=1e-6)


def test_gpu_prototype_fixture(oracle):
    """gpu/fusedmm.cu:60-118: 16 diagonal entries of 2.0 -> out = 2 * mat."""
    g = load_golden("gpu_diag16")
    s, _ = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.SUM)
    assert np.array_equal(s, 2 * g["mat"])


def test_empty_row_convention(oracle):
    """Rows without entries keep the wrapper's init value and the nnz sentinel
    (csrc/fusedmm.cpp:147-150,171): the in-tree evidence, not torch_sparse's 0."""
    g = load_golden("emptyrows_12x9_k5")
    deg = np.diff(g["rowptr"])
    nnz = g["col"].shape[0]
    mx, amx = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.MAX)
    mn, amn = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.MIN)
    assert (mx[deg == 0] == np.finfo(np.float32).min).all() and (amx[deg == 0] == nnz).all()
    assert (mn[deg == 0] == np.finfo(np.float32).max).all() and (amn[deg == 0] == nnz).all()
    assert (amx[deg > 0] < nnz).all() and (amn[deg > 0] < nnz).all()
    s, _ = oracle.spmm_c(g["rowptr"], g["col"], g["value"], g["mat"], oracle.SUM)
    assert (s[deg == 0] == 0).all()


@pytest.mark.parametrize("K", [1, 3, 32, 47, 64, 100, 128])
@pytest.mark.parametrize("with_value", [True, False])
def test_random_graphs_c_vs_numpy(oracle, K, with_value):
    rng = np.random.default_rng(K * 2 + int(with_value))
    M, N = 70, 55
    rowptr, col, val = random_csr(rng, M, N, 40, empty_prob=0.1, with_value=with_value, long_rows=[(3, 300)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    for code in (oracle.SUM, oracle.MEAN, oracle.MAX, oracle.MIN):
        a, aa = oracle.spmm_c(rowptr, col, val, mat, code)
        b, ba = oracle.spmm_numpy(rowptr, col, val, mat, code)
        if code in (oracle.MAX, oracle.MIN):
            assert np.array_equal(a, b) and np.array_equal(aa, ba)
        else:
            # row 3 has 300 entries: sequential fp32 vs pairwise float64 -> condition-aware bound
            assert_sum_close(a, b, abs_product_sum(rowptr, col, val, mat, mean=(code == oracle.MEAN)))


def test_backward_is_the_adjoint(oracle):
    """<A x, g> == <x, A^T g> for sum and mean (float64 check of the CSC-view construction)."""
    rng = np.random.default_rng(7)
    M, N, K = 40, 30, 9
    rowptr, col, val = random_csr(rng, M, N, 25, empty_prob=0.1)
    x = rng.standard_normal((N, K)).astype(np.float32)
    g = rng.standard_normal((M, K)).astype(np.float32)
    for code, bw in ((oracle.SUM, oracle.spmm_backward_sum), (oracle.MEAN, oracle.spmm_backward_mean)):
        y, _ = oracle.spmm_c(rowptr, col, val, x, code)
        gx = bw(rowptr, col, val, g, N)
        lhs = float((y.astype(np.float64) * g).sum())
        rhs = float((x.astype(np.float64) * gx).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def test_build_csc_matches_torch_sparse_definition(oracle):
    rng = np.random.default_rng(11)
    M, N = 33, 21
    rowptr, col, _ = random_csr(rng, M, N, 15, empty_prob=0.2)
    row = np.repeat(np.arange(M), np.diff(rowptr))
    colptr, csr2csc, row_t = oracle.build_csc(rowptr, col, N)
    ref = np.argsort(col * M + row, kind="stable")     # torch_sparse: (col * M + row).argsort()
    assert np.array_equal(csr2csc, ref)
    assert np.array_equal(row_t, row[ref])
    assert np.array_equal(colptr, np.concatenate([[0], np.cumsum(np.bincount(col, minlength=N))]))


def test_unsupported_message_status(oracle):
    import ctypes
    lib = oracle.load_c()
    z = np.zeros((1, 1), np.float32)
    rp = np.zeros(2, np.int64)
    st = lib.fusedMM_csr(0x11101, 1, 1, 1, 1.0, 0, 1, 1, None, None, oracle._ptr(rp),
                         ctypes.c_void_p(rp.ctypes.data + 8), None, 1, None, 1, 0.0, oracle._ptr(z), 1, None)
    assert st == 128   # FUSEDMM_NO_OPT_IMPL, csrc/fusedMM.h:114


def test_oracle_epilogue_matches_the_reference_callers(oracle):
    """The epilogue restatement (oracle.apply_epilogue) against vectors produced by the reference's own
    operator layer + the torch ops its callers apply (tests/golden/make_golden_callers.py)."""
    from conftest import CALLER_CASES, GOLDEN_DIR
    import os
    assert len(CALLER_CASES) >= 2
    for name in CALLER_CASES:
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        val = z["value"] if "value" in z.files else None
        h = (z["x"] @ z["W"].T).astype(np.float32)
        agg = oracle.spmm_c(z["rowptr"], z["col"], val, h, oracle.SUM)[0]
        np.testing.assert_allclose(oracle.apply_epilogue(agg, bias=z["b"], relu=True), z["gcn_out"], rtol=1e-5, atol=1e-5)
        agg = oracle.spmm_c(z["rowptr"], z["col"], None, z["x"], oracle.SUM)[0]
        np.testing.assert_allclose(oracle.apply_epilogue(agg, addend=z["x"], addend_scale=1.0 + float(z["eps"])),
                                   z["gin_out"], rtol=1e-5, atol=1e-5)


def test_compare_conventions_nan_signed_zero_and_ties(oracle):
    """The conventions the reference tree leaves open, pinned for all three restatements:
    strict compare scanning in CSR order -> the smallest edge id wins a tie, +0.0 and -0.0 tie (first seen
    wins), and a NaN product never replaces a candidate (a row of NaNs only keeps the init value + sentinel)."""
    nan = np.float32("nan")
    rowptr = np.array([0, 3, 6, 8, 10], dtype=np.int64)
    col = np.array([0, 1, 2, 0, 1, 2, 3, 3, 4, 5], dtype=np.int64)
    val = np.ones(10, dtype=np.float32)
    #            x[0]  x[1]  x[2]  x[3]  x[4]  x[5]
    mat = np.array([[2.0, 2.0, 1.0, nan, 0.0, -0.0],
                    [nan, 5.0, 5.0, nan, -0.0, 0.0]], dtype=np.float32).T.copy()     # [6, 2]
    nnz = col.shape[0]
    for impl in (oracle.spmm_c, oracle.spmm_loops, oracle.spmm_numpy):
        out, arg = impl(rowptr, col, val, mat, oracle.MAX)
        # row 0, k=0: products 2, 2, 1 -> tie between edges 0 and 1 -> edge 0
        assert out[0, 0] == 2.0 and arg[0, 0] == 0, impl.__name__
        # row 0, k=1: NaN, 5, 5 -> the NaN never wins, first 5 is edge 1
        assert out[0, 1] == 5.0 and arg[0, 1] == 1, impl.__name__
        # row 1 repeats the columns with edge ids 3..5
        assert arg[1, 0] == 3 and arg[1, 1] == 4, impl.__name__
        # row 2: both entries are column 3 (NaN, NaN): nothing ever wins
        assert arg[2, 0] == nnz and arg[2, 1] == nnz, impl.__name__
        assert out[2, 0] == np.finfo(np.float32).min and out[2, 1] == np.finfo(np.float32).min, impl.__name__
        # row 3: +0.0 then -0.0 (k=0), -0.0 then +0.0 (k=1): equal, the first seen (edge 8) stays, bit pattern kept
        assert arg[3, 0] == 8 and arg[3, 1] == 8, impl.__name__
        assert not np.signbit(out[3, 0]) and np.signbit(out[3, 1]), impl.__name__
        out, arg = impl(rowptr, col, val, mat, oracle.MIN)
        assert out[0, 0] == 1.0 and arg[0, 0] == 2 and out[0, 1] == 5.0 and arg[0, 1] == 1, impl.__name__
        assert arg[2, 0] == nnz and out[2, 0] == np.finfo(np.float32).max, impl.__name__
        assert arg[3, 0] == 8 and arg[3, 1] == 8, impl.__name__

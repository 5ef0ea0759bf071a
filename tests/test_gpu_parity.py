"""GPU parity tests (run with -m gpu on the B200 box): the CUDA kernels, called THROUGH
THE C ABI (isplib_b200/capi.py -> libisplib_b200.so), against the CPU oracle on the same
seeded inputs, against the committed golden vectors, and -- at full benchmark size --
through size-independent properties.

Bar: max/min out and arg bit-exact; sum/mean within rtol 1e-5 /
atol 1e-6 (condition-aware where a row's terms cancel, see conftest.assert_sum_close).
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, abs_product_sum, assert_sum_close, grad_cond, load_golden, random_csr

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
REDUCES = ("sum", "mean", "max", "min")


@pytest.fixture(scope="module")
def capi():
    from isplib_b200 import capi as c
    c.lib()
    return c


def to_dev(rowptr, col, val, mat):
    rp = torch.from_numpy(np.ascontiguousarray(rowptr)).to(DEV).to(torch.int32)
    co = torch.from_numpy(np.ascontiguousarray(col)).to(DEV).to(torch.int32)
    va = None if val is None else torch.from_numpy(np.ascontiguousarray(val)).to(DEV)
    x = torch.from_numpy(np.ascontiguousarray(mat)).to(DEV)
    return rp, co, va, x


def check_forward(capi, oracle, rowptr, col, val, mat, reduce, variant=-1, seg_len=0, plan=None):
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = plan or capi.Plan(rp, co.numel(), seg_len)
    out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, variant)
    torch.cuda.synchronize()
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    if reduce in ("max", "min"):
        assert np.array_equal(out.cpu().numpy(), ref), f"{reduce} out not bit-exact"
        assert np.array_equal(arg.cpu().numpy(), ref_arg), f"{reduce} arg not bit-exact"
    else:
        cond = abs_product_sum(rowptr, col, val, mat, mean=(reduce == "mean"))
        assert_sum_close(out.cpu().numpy(), ref, cond)
    return out, arg


# ----------------------------------------------------------------------------- golden
@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("reduce", REDUCES)
def test_golden_forward(capi, oracle, name, reduce):
    g = load_golden(name)
    rp, co, va, x = to_dev(g["rowptr"], g["col"], g["value"], g["mat"])
    plan = capi.Plan(rp, co.numel())
    out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan)
    if reduce in ("max", "min"):
        assert np.array_equal(out.cpu().numpy(), g[f"{reduce}_out"])
        assert np.array_equal(arg.cpu().numpy(), g[f"{reduce}_arg"])
    else:
        # the golden vectors come from the reference's own operator layer: here the PLAIN north_star
        # tolerance is asserted (rows are short; no element needs the condition-aware bound)
        np.testing.assert_allclose(out.cpu().numpy(), g[f"{reduce}_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_backward(capi, oracle, name):
    """sum/mean backward = forward kernel on the device-built CSC view; max/min backward =
    fused arg scatter; both against the reference autograd's gradients."""
    g = load_golden(name)
    N = g["N"]
    rp, co, va, x = to_dev(g["rowptr"], g["col"], g["value"], g["mat"])
    go = torch.from_numpy(g["grad_out"]).to(DEV)
    colptr, row_t, csr2csc = capi.csr_transpose(rp, co, N)
    plan_t = capi.Plan(colptr, co.numel())
    vt = capi.permute_values(va, csr2csc, row_t, rp, False) if va is not None else None
    gs, _ = capi.spmm_csr("sum", colptr, row_t, vt, go, plan_t)
    assert_sum_close(gs.cpu().numpy(), g["sum_grad_mat"], grad_cond(oracle, g["rowptr"], g["col"], g["value"], g["grad_out"], N))
    w = capi.permute_values(va, csr2csc, row_t, rp, True)
    gm, _ = capi.spmm_csr("sum", colptr, row_t, w, go, plan_t)
    assert_sum_close(gm.cpu().numpy(), g["mean_grad_mat"],
                     grad_cond(oracle, g["rowptr"], g["col"], g["value"], g["grad_out"], N, mean=True))
    for red in ("max", "min"):
        arg = torch.from_numpy(g[f"{red}_arg"]).to(DEV)
        val = va if va is not None else torch.ones(co.numel(), device=DEV)
        gx, gv = capi.spmm_arg_backward(co, val, x, arg, go, N, True, True)
        np.testing.assert_allclose(gx.cpu().numpy(), g[f"{red}_grad_mat"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gv.cpu().numpy(), g[f"{red}_grad_value"], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------- seeded random graphs
@pytest.mark.parametrize("K", [1, 3, 4, 8, 16, 32, 47, 64, 100, 128, 200, 256, 260, 602])
@pytest.mark.parametrize("reduce", REDUCES)
def test_random_graph_all_k(capi, oracle, K, reduce):
    rng = np.random.default_rng(1000 + K)
    M, N = 300, 257
    rowptr, col, val = random_csr(rng, M, N, 70, empty_prob=0.05, long_rows=[(5, 1500), (17, 257), (18, 256)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    check_forward(capi, oracle, rowptr, col, val, mat, reduce)


@pytest.mark.parametrize("reduce", REDUCES)
def test_every_variant(capi, oracle, reduce):
    rng = np.random.default_rng(5)
    M, N, K = 500, 400, 128
    rowptr, col, val = random_csr(rng, M, N, 90, empty_prob=0.02, long_rows=[(0, 3000), (499, 700)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = capi.Plan(rp, co.numel())
    L = capi.lib()
    ran = 0
    for v in range(L.isplib_b200_variant_count()):
        if not L.isplib_b200_variant_supported(v, capi.REDUCE_CODE[reduce], K, K, K, x.data_ptr(), x.data_ptr()):
            continue
        check_forward(capi, oracle, rowptr, col, val, mat, reduce, variant=v, plan=plan)
        ran += 1
    assert ran >= 4


@pytest.mark.parametrize("K", [32, 40, 64, 104, 200, 256])   # 40/104/200: ragged tiles of the lean 256-bit kernel
def test_every_variant_other_widths(capi, oracle, K):
    rng = np.random.default_rng(6 + K)
    M, N = 200, 300
    rowptr, col, val = random_csr(rng, M, N, 60, long_rows=[(7, 900)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    plan = capi.Plan(rp, co.numel())
    L = capi.lib()
    for v in range(L.isplib_b200_variant_count()):
        for reduce in ("sum", "max"):
            if L.isplib_b200_variant_supported(v, capi.REDUCE_CODE[reduce], K, K, K, x.data_ptr(), x.data_ptr()):
                check_forward(capi, oracle, rowptr, col, val, mat, reduce, variant=v, plan=plan)


@pytest.mark.parametrize("seg_len", [32, 64, 128, 512, 1024])
@pytest.mark.parametrize("reduce", ["sum", "max", "min"])
def test_segment_lengths(capi, oracle, seg_len, reduce):
    """However a row is split into segments, max/min/arg stay bit-exact."""
    rng = np.random.default_rng(seg_len)
    M, N, K = 64, 90, 36
    rowptr, col, val = random_csr(rng, M, N, 50, long_rows=[(1, 2049), (2, seg_len), (3, seg_len + 1), (4, 2 * seg_len)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    check_forward(capi, oracle, rowptr, col, val, mat, reduce, seg_len=seg_len)


@pytest.mark.parametrize("reduce", REDUCES)
def test_no_value_matrix(capi, oracle, reduce):
    rng = np.random.default_rng(9)
    rowptr, col, _ = random_csr(rng, 128, 128, 40, with_value=False, long_rows=[(9, 600)])
    mat = rng.standard_normal((128, 64)).astype(np.float32)
    check_forward(capi, oracle, rowptr, col, None, mat, reduce)


@pytest.mark.parametrize("reduce", ["max", "min"])
def test_ties_and_signed_zeros(capi, oracle, reduce):
    """Integer-valued products create thousands of exact ties (and +0/-0 pairs): the
    smallest edge id must win, also across lane groups (K=32 -> 4 entries per step) and
    across segments (row 0 has 1000 entries)."""
    rng = np.random.default_rng(3)
    for K in (8, 32, 64, 128):
        M, N = 40, 12
        rowptr, col, _ = random_csr(rng, M, N, 64, long_rows=[(0, 1000)])
        val = rng.integers(-2, 3, size=col.shape[0]).astype(np.float32)
        mat = rng.integers(-2, 3, size=(N, K)).astype(np.float32)
        check_forward(capi, oracle, rowptr, col, val, mat, reduce)


def test_empty_graph_and_all_empty_rows(capi, oracle):
    for M, N, K in ((5, 4, 8), (1, 1, 1)):
        rowptr = np.zeros(M + 1, dtype=np.int64)
        col = np.zeros(0, dtype=np.int64)
        val = np.zeros(0, dtype=np.float32)
        mat = np.ones((N, K), dtype=np.float32)
        for reduce in REDUCES:
            out, arg = check_forward(capi, oracle, rowptr, col, val, mat, reduce)
            if reduce == "max":
                assert (out == torch.finfo(torch.float32).min).all() and (arg == 0).all()


def test_empty_zero_flag(capi, oracle):
    """ISPLIB_FLAG_EMPTY_ZERO: torch_sparse's convention (0 in rows without entries)."""
    g = load_golden("emptyrows_12x9_k5")
    rp, co, va, x = to_dev(g["rowptr"], g["col"], g["value"], g["mat"])
    plan = capi.Plan(rp, co.numel())
    deg = np.diff(g["rowptr"])
    for red in ("max", "min"):
        out, arg = capi.spmm_csr(red, rp, co, va, x, plan, flags=capi.FLAG_EMPTY_ZERO)
        o = out.cpu().numpy()
        assert (o[deg == 0] == 0).all()
        assert np.array_equal(o[deg > 0], g[f"{red}_out"][deg > 0])
        assert np.array_equal(arg.cpu().numpy(), g[f"{red}_arg"])


def test_strided_operands(capi, oracle):
    """ldx / ldo larger than K (views into wider matrices)."""
    rng = np.random.default_rng(21)
    M, N, K = 100, 80, 64
    rowptr, col, val = random_csr(rng, M, N, 30)
    big = rng.standard_normal((N, K + 32)).astype(np.float32)
    rp, co, va, xb = to_dev(rowptr, col, val, big)
    x = xb[:, 16:16 + K]
    outb = torch.full((M, K + 8), 7.0, device=DEV)
    plan = capi.Plan(rp, co.numel())
    for reduce in ("sum", "max"):
        out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, out=outb[:, 4:4 + K])
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, np.ascontiguousarray(big[:, 16:16 + K]), oracle.REDUCE_CODE[reduce])
        if reduce == "max":
            assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg)
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, big[:, 16:16 + K]))
        assert (outb[:, :4] == 7.0).all() and (outb[:, 4 + K:] == 7.0).all()


def test_accumulate_and_edge_ids_column_blocks(capi, oracle):
    """The multi-GPU building block: A split by column range into two CSR blocks, second
    call ACCUMULATEs; max/min use global edge ids so the result equals the 1-block answer."""
    rng = np.random.default_rng(33)
    M, N, K = 90, 100, 32
    rowptr, col, val = random_csr(rng, M, N, 50, empty_prob=0.1, long_rows=[(4, 800)])
    # make ties likely for max/min
    val = np.round(val * 2).astype(np.float32)
    mat = rng.integers(-3, 4, size=(N, K)).astype(np.float32)
    row = np.repeat(np.arange(M), np.diff(rowptr))
    eid = np.arange(col.shape[0])
    blocks = []
    for lo, hi in ((0, 37), (37, N)):
        sel = (col >= lo) & (col < hi)
        rp = np.zeros(M + 1, dtype=np.int64)
        rp[1:] = np.cumsum(np.bincount(row[sel], minlength=M))
        blocks.append((rp, col[sel] - lo, val[sel], eid[sel], mat[lo:hi]))
    nnz = col.shape[0]
    deg = torch.from_numpy(np.maximum(np.diff(rowptr), 1).astype(np.float32)).to(DEV)
    for reduce in REDUCES:
        out = arg = None
        for bi, (rp, c, v, e, x) in enumerate(blocks):
            rpd, cd, vd, xd = to_dev(rp, c, v, x)
            ed = torch.from_numpy(e).to(DEV).to(torch.int32)
            plan = capi.Plan(rpd, cd.numel())
            last = bi == len(blocks) - 1
            out, arg = capi.spmm_csr("sum" if reduce == "mean" else reduce, rpd, cd, vd, xd, plan, out=out, arg_out=arg,
                                     flags=capi.FLAG_ACCUMULATE if bi else 0,
                                     row_divisor=deg if (reduce == "mean" and last) else None,
                                     edge_ids=ed, arg_sentinel=nnz)
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
        if reduce in ("max", "min"):
            assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg)
        else:
            assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, reduce == "mean"))


# ----------------------------------------------------------------- graph ops
def test_csr_transpose_matches_oracle(capi, oracle):
    rng = np.random.default_rng(44)
    for M, N in ((50, 70), (300, 40), (1, 5), (7, 1)):
        rowptr, col, val = random_csr(rng, M, N, 25, empty_prob=0.2)
        rp, co, va, _ = to_dev(rowptr, col, val, np.zeros((1, 1), np.float32))
        colptr, row_t, csr2csc = capi.csr_transpose(rp, co, N)
        rc, rs, rr = oracle.build_csc(rowptr, col, N)
        assert np.array_equal(colptr.cpu().numpy(), rc)
        assert np.array_equal(csr2csc.cpu().numpy(), rs)
        assert np.array_equal(row_t.cpu().numpy(), rr)
        w = capi.permute_values(va, csr2csc, row_t, rp, True).cpu().numpy()
        deg = np.maximum(np.diff(rowptr), 1).astype(np.float32)
        assert np.array_equal(w, val[rs] / deg[rr])


def test_plan_info(capi):
    deg = np.array([0, 1, 256, 257, 1000, 0, 3], dtype=np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    rp = torch.from_numpy(rowptr).to(DEV).to(torch.int32)
    plan = capi.Plan(rp, int(rowptr[-1]), 256)
    i = plan.info
    assert (i.m, i.nnz, i.seg_len) == (7, int(rowptr[-1]), 256)
    assert i.num_items == 1 + 1 + 1 + 2 + 4 + 1 + 1
    assert i.num_split_rows == 2 and i.num_split_items == 6
    assert i.max_degree == 1000 and i.num_empty_rows == 2


def test_narrow_overflow_detected(capi):
    ok = torch.tensor([0, 5, 2**31 - 1], dtype=torch.int64, device=DEV)
    assert capi.narrow_i64_to_i32(ok).tolist() == [0, 5, 2**31 - 1]
    with pytest.raises(capi.IsplibError):
        capi.narrow_i64_to_i32(torch.tensor([0, 2**31], dtype=torch.int64, device=DEV))


def test_bad_arguments_return_status(capi):
    rp = torch.tensor([0, 1], dtype=torch.int32, device=DEV)
    co = torch.tensor([0], dtype=torch.int32, device=DEV)
    x = torch.ones((1, 4), device=DEV)
    plan = capi.Plan(rp, 1)
    with pytest.raises(capi.IsplibError) as e:
        capi.spmm_csr(7, rp, co, None, x, plan)            # unknown reduction
    assert e.value.status == 128
    with pytest.raises(capi.IsplibError) as e:
        capi.spmm_csr("sum", rp, co, None, x, plan, variant=999)
    assert e.value.status == 128
    import ctypes
    out = torch.empty((1, 4), device=DEV)
    L = capi.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    # plan built for m=1 used with m=2: FUSEDMM_FAIL_RETURN, nothing launched
    st = L.isplib_b200_spmm_csr(capi.SUM, 2, 1, 4, 1, P(rp), P(co), None, P(x), 4, P(out), 4, None,
                                ctypes.byref(plan.info), plan.ptr, None, 0, -1, None)
    assert st == 1
    # max without arg_out
    st = L.isplib_b200_spmm_csr(capi.MAX, 1, 1, 4, 1, P(rp), P(co), None, P(x), 4, P(out), 4, None,
                                ctypes.byref(plan.info), plan.ptr, None, 0, -1, None)
    assert st == 256
    # a split row needs workspace: none given -> FUSEDMM_NOT_ENOUGH_MEM
    rp2 = torch.tensor([0, 600], dtype=torch.int32, device=DEV)
    co2 = torch.zeros(600, dtype=torch.int32, device=DEV)
    plan2 = capi.Plan(rp2, 600, 256)
    st = L.isplib_b200_spmm_csr(capi.SUM, 1, 1, 4, 600, P(rp2), P(co2), None, P(x), 4, P(out), 4, None,
                                ctypes.byref(plan2.info), plan2.ptr, None, 0, -1, None)
    assert st == -1


# ----------------------------------------------------------------- arg backward
@pytest.mark.parametrize("reduce", ["max", "min"])
def test_arg_backward_random(capi, oracle, reduce):
    rng = np.random.default_rng(55)
    M, N, K = 200, 150, 48
    rowptr, col, val = random_csr(rng, M, N, 40, empty_prob=0.1)
    mat = rng.standard_normal((N, K)).astype(np.float32)
    go = rng.standard_normal((M, K)).astype(np.float32)
    _, arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    rgx, rgv = oracle.arg_backward(col, val, mat, arg, go, N, True)
    rp, co, va, x = to_dev(rowptr, col, val, mat)
    gx, gv = capi.spmm_arg_backward(co, va, x, torch.from_numpy(arg).to(DEV), torch.from_numpy(go).to(DEV), N, True, True)
    np.testing.assert_allclose(gx.cpu().numpy(), rgx, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gv.cpu().numpy(), rgv, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------- reference entry point
@pytest.mark.parametrize("reduce", REDUCES)
def test_fusedmm_csr_host_entry(capi, oracle, reduce):
    """isplib_b200_fusedmm_csr_host has the reference's exact fusedMM_csr contract
    (csrc/fusedMM.h:77-99): host pointers, int64, accumulate into pre-initialised z."""
    import ctypes
    rng = np.random.default_rng(66)
    M, N, K = 60, 50, 24
    rowptr, col, val = random_csr(rng, M, N, 30, empty_prob=0.1, long_rows=[(2, 600)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    code = oracle.REDUCE_CODE[reduce]
    out, arg = oracle._init_out(M, K, col.shape[0], code)
    dummy = np.zeros(1, np.float32)
    st = capi.lib().isplib_b200_fusedmm_csr_host(
        oracle.IMSG[code], M, N, K, 1.0, col.shape[0], M, N, oracle._ptr(val), oracle._ptr(col), oracle._ptr(rowptr),
        ctypes.c_void_p(rowptr.ctypes.data + 8), oracle._ptr(dummy), K, oracle._ptr(mat), K, 0.0, oracle._ptr(out), K,
        oracle._ptr(arg))
    assert st == 0
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, code)
    if reduce in ("max", "min"):
        assert np.array_equal(out, ref) and np.array_equal(arg, ref_arg)
    else:
        assert_sum_close(out, ref, abs_product_sum(rowptr, col, val, mat, reduce == "mean"))


def test_reference_gpu_prototype_kernel_agrees(capi, oracle):
    """The only SpMM arithmetic actually in the reference tree, gpu/kernels/spmm.cuh:3-23
    (sum, int64 indices), compiled from where it lies into oracle/_ref/libref_gpu_proto.so."""
    import ctypes, os
    if not os.path.exists(oracle.REF_GPU_PROTO_PATH):
        pytest.skip("oracle/_ref/libref_gpu_proto.so not built (needs /root/reference at build time)")
    ref_lib = ctypes.CDLL(oracle.REF_GPU_PROTO_PATH)
    rng = np.random.default_rng(77)
    M, N, K = 120, 90, 40
    rowptr, col, val = random_csr(rng, M, N, 35)
    mat = rng.standard_normal((N, K)).astype(np.float32)
    rp64 = torch.from_numpy(rowptr).to(DEV)
    co64 = torch.from_numpy(col).to(DEV)
    va = torch.from_numpy(val).to(DEV)
    x = torch.from_numpy(mat).to(DEV)
    c = torch.zeros((M, K), device=DEV)
    torch.cuda.synchronize()
    err = ref_lib.ref_gpu_proto_spmm_sum(M, N, K, col.shape[0], ctypes.c_void_p(co64.data_ptr()),
                                         ctypes.c_void_p(rp64.data_ptr()), ctypes.c_void_p(va.data_ptr()),
                                         ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(c.data_ptr()))
    torch.cuda.synchronize()
    assert err == 0
    ref, _ = oracle.spmm_c(rowptr, col, val, mat, oracle.SUM)
    cond = abs_product_sum(rowptr, col, val, mat)
    assert_sum_close(c.cpu().numpy(), ref, cond)               # reference GPU kernel vs oracle
    plan = capi.Plan(rp64.to(torch.int32), col.shape[0])
    out, _ = capi.spmm_csr("sum", rp64.to(torch.int32), co64.to(torch.int32), va, x, plan)
    assert_sum_close(out.cpu().numpy(), c.cpu().numpy(), cond)  # ours vs reference GPU kernel


@pytest.mark.parametrize("K", [5, 47, 101, 602])
@pytest.mark.parametrize("reduce", REDUCES)
def test_padded_rows_vector_gather_odd_k(capi, oracle, K, reduce):
    """K % 4 != 0 with x rows padded to a multiple of 4 floats (what the op layer does for
    widths like 47 / 602): 16-byte gathers may read the padding, stores must not go past K."""
    rng = np.random.default_rng(300 + K)
    M, N = 150, 120
    rowptr, col, val = random_csr(rng, M, N, 40, empty_prob=0.05, long_rows=[(2, 1300)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    Kp = (K + 3) // 4 * 4
    rp, co, va, _ = to_dev(rowptr, col, val, mat)
    xpad = torch.full((N, Kp), float("nan"), device=DEV)     # NaN padding must never leak
    xpad[:, :K] = torch.from_numpy(mat).to(DEV)
    x = xpad[:, :K]
    outbuf = torch.full((M, K + 3), 7.0, device=DEV)
    plan = capi.Plan(rp, co.numel())
    out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, out=outbuf[:, :K])
    ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, oracle.REDUCE_CODE[reduce])
    if reduce in ("max", "min"):
        assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg)
    else:
        assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat, reduce == "mean"))
    assert (outbuf[:, K:] == 7.0).all()


# ----------------------------------------------------------------- SDDMM (grad_value of sum/mean)
@pytest.mark.parametrize("K", [1, 4, 7, 32, 40, 47, 64, 100, 128, 200, 256, 300, 1030])
@pytest.mark.parametrize("mean", [False, True])
def test_sddmm_matches_oracle(capi, oracle, K, mean):
    rng = np.random.default_rng(700 + K)
    M, N = 120, 90
    rowptr, col, _ = random_csr(rng, M, N, 45, empty_prob=0.1, with_value=False, long_rows=[(4, 1100), (7, 33)])
    a = rng.standard_normal((M, K)).astype(np.float32)
    x = rng.standard_normal((N, K)).astype(np.float32)
    rp, co, _, xd = to_dev(rowptr, col, None, x)
    ad = torch.from_numpy(a).to(DEV)
    plan = capi.Plan(rp, co.numel())
    got = capi.sddmm_csr(rp, co, ad, xd, plan, mean).cpu().numpy()
    ref = oracle.sddmm(rowptr, col, a, x, mean)
    row = np.repeat(np.arange(M), np.diff(rowptr))
    cond = (np.abs(a.astype(np.float64))[row] * np.abs(x.astype(np.float64))[col]).sum(axis=1)
    if mean:
        cond = cond / np.maximum(np.diff(rowptr), 1)[row]
    assert_sum_close(got, ref, cond)
    # padded rows (what the op layer hands over for odd K): NaN padding must not leak into the dot;
    # rows padded to 8 floats are 32-byte aligned, which is what selects the 256-bit kernel
    for pad in (4, 8):
        if K % pad:
            Kp = (K + pad - 1) // pad * pad
            xp = torch.full((N, Kp), float("nan"), device=DEV)
            xp[:, :K] = xd
            ap = torch.full((M, Kp), float("nan"), device=DEV)
            ap[:, :K] = ad
            got2 = capi.sddmm_csr(rp, co, ap[:, :K], xp[:, :K], plan, mean).cpu().numpy()
            assert_sum_close(got2, ref, cond)


@pytest.mark.parametrize("pad", [4, 8])   # 8: rows stay 32-byte aligned, so the 256-bit kernels run too
@pytest.mark.parametrize("K", [40, 47, 64, 100, 128, 200])
def test_no_out_of_bounds_writes_canaries(capi, oracle, K, pad):
    """compute-sanitizer is not available on this pool, so out-of-bounds WRITES are caught with
    canaries: out / arg_out live inside larger buffers whose guard rows and guard columns must
    be untouched after every variant has run (split rows, partial tiles, scalar-store tail)."""
    rng = np.random.default_rng(900 + K)
    M, N = 97, 80
    rowptr, col, val = random_csr(rng, M, N, 40, empty_prob=0.1, long_rows=[(0, 1200), (96, 700)])
    mat = rng.standard_normal((N, K)).astype(np.float32)
    rp, co, va, _ = to_dev(rowptr, col, val, mat)
    Kp = (K + pad - 1) // pad * pad
    xb = torch.full((N + 2, Kp + pad), 3.0, device=DEV)
    xb[1:N + 1, :K] = torch.from_numpy(mat).to(DEV)
    x = xb[1:N + 1, :K]
    plan = capi.Plan(rp, co.numel(), 256)
    L = capi.lib()
    for reduce in ("sum", "max"):
        code = capi.REDUCE_CODE[reduce]
        ref, ref_arg = oracle.spmm_c(rowptr, col, val, mat, code)
        for v in range(L.isplib_b200_variant_count()):
            if not L.isplib_b200_variant_supported(v, code, K, x.stride(0), Kp + pad, x.data_ptr(), x.data_ptr()):
                continue
            ob = torch.full((M + 2, Kp + pad), 7.0, device=DEV)
            ab = torch.full((M + 2, Kp + pad), -5, dtype=torch.int64, device=DEV)
            out, arg = capi.spmm_csr(reduce, rp, co, va, x, plan, v, out=ob[1:M + 1, :K],
                                     arg_out=ab[1:M + 1, :K] if reduce == "max" else None)
            torch.cuda.synchronize()
            name = capi.variant_names()[v]
            assert (ob[0] == 7).all() and (ob[M + 1] == 7).all() and (ob[:, K:] == 7).all(), name
            if reduce == "max":
                assert (ab[0] == -5).all() and (ab[M + 1] == -5).all() and (ab[:, K:] == -5).all(), name
                assert np.array_equal(out.cpu().numpy(), ref) and np.array_equal(arg.cpu().numpy(), ref_arg), name
            else:
                assert_sum_close(out.cpu().numpy(), ref, abs_product_sum(rowptr, col, val, mat))


def test_l2_tiled_paths_sddmm_and_arg_backward(capi, oracle):
    """The K-tiled code paths only switch on when X / grad_x exceed the L2 budget (> 64-96 MB):
    a wide dense operand (200k x 128 = 102 MB) with a small graph on top exercises the tiled
    SDDMM (two 64-wide launches accumulating) and the tiled arg scatter (32-wide slabs)."""
    rng = np.random.default_rng(4242)
    M, N, K = 300, 200_000, 128
    rowptr, col, val = random_csr(rng, M, N, 60, empty_prob=0.05, long_rows=[(1, 900)])
    x = rng.standard_normal((N, K)).astype(np.float32)
    a = rng.standard_normal((M, K)).astype(np.float32)
    rp, co, va, xd = to_dev(rowptr, col, val, x)
    ad = torch.from_numpy(a).to(DEV)
    plan = capi.Plan(rp, co.numel())
    row = np.repeat(np.arange(M), np.diff(rowptr))
    cond = (np.abs(a.astype(np.float64))[row] * np.abs(x.astype(np.float64))[col]).sum(axis=1)
    for mean in (False, True):
        got = capi.sddmm_csr(rp, co, ad, xd, plan, mean).cpu().numpy()
        ref = oracle.sddmm(rowptr, col, a, x, mean)
        assert_sum_close(got, ref, cond / (np.maximum(np.diff(rowptr), 1)[row] if mean else 1.0))
    _, arg = oracle.spmm_c(rowptr, col, val, x, oracle.MAX)
    rgx, _ = oracle.arg_backward(col, val, None, arg, a, N)
    gx, _ = capi.spmm_arg_backward(co, va, None, torch.from_numpy(arg).to(DEV), ad, N, True, False)
    np.testing.assert_allclose(gx.cpu().numpy(), rgx, rtol=1e-5, atol=1e-5)

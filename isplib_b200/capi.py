"""ctypes binding of the C ABI in ``include/isplib_b200.h``.

This is the host-side mirror used by the parity tests (which call the kernels
through the C ABI, not through torch ops) and by the multi-GPU layer.  Tensors are
only used as owners of device memory: every call passes raw ``data_ptr()`` values,
sizes and the current CUDA stream.  Nothing here falls back to the CPU -- if the
library is missing, importing it raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISPLIB_B200_LIB") or os.path.join(_HERE, "libisplib_b200.so")   # env: A/B builds in tools/

SUM, MAX, MIN, MEAN = 0, 1, 2, 3
REDUCE_CODE = {"sum": SUM, "add": SUM, "max": MAX, "min": MIN, "mean": MEAN}
FLAG_ACCUMULATE = 0x1
FLAG_EMPTY_ZERO = 0x2
FLAG_RELU = 0x4
VARIANT_AUTO = -1

# every symbol include/isplib_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "isplib_b200_abi_version", "isplib_b200_status_string",
    "isplib_b200_plan_bytes", "isplib_b200_plan_build", "isplib_b200_spmm_workspace_bytes",
    "isplib_b200_spmm_csr", "isplib_b200_spmm_csr_ex", "isplib_b200_spmm_csr_fused",
    "isplib_b200_spmm_arg_backward_aux", "isplib_b200_spmm_csr_gather",
    "isplib_b200_spmm_arg_backward_binned", "isplib_b200_spmm_arg_backward_binned_workspace_bytes",
    "isplib_b200_plan_grouped_bytes", "isplib_b200_plan_build_grouped",
    "isplib_b200_variant_count", "isplib_b200_variant_name", "isplib_b200_variant_supported",
    "isplib_b200_variant_default", "isplib_b200_spmm_autotune",
    "isplib_b200_csr_transpose_workspace_bytes", "isplib_b200_csr_transpose",
    "isplib_b200_permute_values", "isplib_b200_spmm_arg_backward",
    "isplib_b200_narrow_i64_to_i32", "isplib_b200_fusedmm_csr_host", "isplib_b200_sddmm_csr",
    "isplib_b200_coo_to_csr_workspace_bytes", "isplib_b200_coo_to_csr",
]


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("m", ctypes.c_int64), ("nnz", ctypes.c_int64),
        ("seg_len", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("num_items", ctypes.c_int64), ("num_split_rows", ctypes.c_int64),
        ("num_split_items", ctypes.c_int64), ("max_degree", ctypes.c_int64),
        ("num_empty_rows", ctypes.c_int64), ("plan_bytes", ctypes.c_uint64),
    ]


class Epilogue(ctypes.Structure):
    """isplib_b200_epilogue (include/isplib_b200.h)."""
    _fields_ = [
        ("bias", ctypes.c_void_p), ("addend", ctypes.c_void_p), ("ld_addend", ctypes.c_int64),
        ("addend_scale", ctypes.c_float), ("reserved", ctypes.c_int32),
        ("arg_col", ctypes.c_void_p), ("arg_val", ctypes.c_void_p),
    ]


class GatherDesc(ctypes.Structure):
    """isplib_b200_gather_desc (include/isplib_b200.h)."""
    _fields_ = [
        ("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("n_groups", ctypes.c_int32), ("copy_ctas", ctypes.c_int32),
        ("peer_x", ctypes.POINTER(ctypes.c_void_p)), ("peer_arrive", ctypes.POINTER(ctypes.c_void_p)),
        ("peer_credit", ctypes.POINTER(ctypes.c_void_p)),
        ("owner_group", ctypes.POINTER(ctypes.c_int32)), ("my_group_at_peer", ctypes.POINTER(ctypes.c_int32)),
        ("slice_rows", ctypes.c_int64), ("status", ctypes.c_void_p), ("epoch", ctypes.c_uint32),
        ("tile_mode", ctypes.c_uint32), ("group_item_end", ctypes.POINTER(ctypes.c_int64)),
        ("parity_launch", ctypes.c_uint32), ("phase", ctypes.c_uint32),
    ]


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"isplib_b200: {LIB_PATH} is missing -- build it with `make -C isplib_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    i32, i64, f32, p, sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
    pinfo = ctypes.POINTER(PlanInfo)
    L.isplib_b200_abi_version.restype = ctypes.c_int
    L.isplib_b200_status_string.restype = ctypes.c_char_p
    L.isplib_b200_status_string.argtypes = [ctypes.c_int]
    L.isplib_b200_plan_bytes.argtypes = [i64, i64, i32, ctypes.POINTER(sz)]
    L.isplib_b200_plan_build.argtypes = [i64, i64, p, i32, p, sz, pinfo, p]
    L.isplib_b200_spmm_workspace_bytes.argtypes = [pinfo, i64, ctypes.c_int, ctypes.POINTER(sz)]
    spmm_args = [ctypes.c_int, i64, i64, i64, i64, p, p, p, p, i64, p, i64, p, pinfo, p, p, sz]
    L.isplib_b200_spmm_csr.argtypes = spmm_args + [ctypes.c_int, p]
    L.isplib_b200_spmm_csr_ex.argtypes = spmm_args + [ctypes.c_int, ctypes.c_int, p, p, i64, p]
    L.isplib_b200_spmm_csr_fused.argtypes = spmm_args + [ctypes.c_int, ctypes.c_int, p, p, i64, ctypes.POINTER(Epilogue), p]
    L.isplib_b200_spmm_csr_gather.argtypes = spmm_args + [ctypes.c_int, ctypes.c_int, p, p, i64, ctypes.POINTER(Epilogue),
                                                          ctypes.POINTER(GatherDesc), p]
    L.isplib_b200_plan_grouped_bytes.argtypes = [i64, i64, i32, i32, ctypes.POINTER(sz)]
    L.isplib_b200_plan_build_grouped.argtypes = [i64, i64, p, p, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), i32,
                                                 p, sz, pinfo, ctypes.POINTER(i64), p]
    L.isplib_b200_spmm_arg_backward_binned_workspace_bytes.argtypes = [i64, i64, i64, ctypes.POINTER(sz)]
    L.isplib_b200_spmm_arg_backward_binned.argtypes = [i64, i64, i64, p, p, i64, p, i64, p, i64, ctypes.c_int, p, sz, p]
    L.isplib_b200_spmm_arg_backward_aux.argtypes = [i64, i64, i64, p, p, i64, p, i64, p, i64, ctypes.c_int, p]
    L.isplib_b200_variant_count.restype = ctypes.c_int
    L.isplib_b200_variant_name.restype = ctypes.c_char_p
    L.isplib_b200_variant_name.argtypes = [ctypes.c_int]
    L.isplib_b200_variant_supported.argtypes = [ctypes.c_int, ctypes.c_int, i64, i64, i64, p, p]
    L.isplib_b200_variant_default.argtypes = [ctypes.c_int, i64, i64, i64, i64, p, p, ctypes.c_double]
    L.isplib_b200_spmm_autotune.argtypes = spmm_args + [ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                        ctypes.POINTER(f32), p]
    L.isplib_b200_csr_transpose_workspace_bytes.argtypes = [i64, i64, i64, ctypes.POINTER(sz)]
    L.isplib_b200_csr_transpose.argtypes = [i64, i64, i64, p, p, p, p, p, p, sz, p]
    L.isplib_b200_permute_values.argtypes = [i64, p, p, p, p, ctypes.c_int, p, p]
    L.isplib_b200_spmm_arg_backward.argtypes = [i64, i64, i64, i64, p, p, p, i64, p, i64, i64, p, i64,
                                                p, i64, p, ctypes.c_int, p]
    L.isplib_b200_narrow_i64_to_i32.argtypes = [i64, p, p, p, p]
    L.isplib_b200_coo_to_csr_workspace_bytes.argtypes = [i64, i64, i64, ctypes.POINTER(sz)]
    L.isplib_b200_coo_to_csr.argtypes = [i64, i64, i64, p, p, p, p, p, p, p, p, sz, p]
    L.isplib_b200_sddmm_csr.argtypes = [i64, i64, i64, i64, p, p, p, i64, p, i64, ctypes.c_int, p, pinfo, p, p]
    L.isplib_b200_fusedmm_csr_host.argtypes = [i32, i64, i64, i64, f32, i64, i64, i64, p, p, p, p, p, i64,
                                               p, i64, f32, p, i64, p]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError here = the .so does not match include/isplib_b200.h
    _lib = L
    return L


class IsplibError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().isplib_b200_status_string(status).decode()
        super().__init__(f"{where} failed: {msg} (status {status})")


def check(status: int, where: str) -> None:
    if status != 0:
        raise IsplibError(status, where)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_bytes(nbytes: int, device) -> torch.Tensor:
    # +256 so the 256-byte aligned window of `nbytes` always fits
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


def _aligned_ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p((t.data_ptr() + 255) // 256 * 256)


class Plan:
    """Owner of a device plan buffer + its host-side info (isplib_b200_plan_build)."""

    def __init__(self, rowptr32: torch.Tensor, nnz: int, seg_len: int = 0):
        assert rowptr32.is_cuda and rowptr32.dtype == torch.int32 and rowptr32.is_contiguous()
        self.m = rowptr32.numel() - 1
        self.nnz = int(nnz)
        nbytes = ctypes.c_size_t(0)
        check(lib().isplib_b200_plan_bytes(self.m, self.nnz, seg_len, ctypes.byref(nbytes)), "plan_bytes")
        self.buf = _dev_bytes(nbytes.value, rowptr32.device)
        self.info = PlanInfo()
        check(lib().isplib_b200_plan_build(self.m, self.nnz, _p(rowptr32), seg_len, _aligned_ptr(self.buf),
                                           nbytes.value, ctypes.byref(self.info), _stream(rowptr32.device)),
              "plan_build")

    @property
    def ptr(self) -> ctypes.c_void_p:
        return _aligned_ptr(self.buf)


class GroupedPlan(Plan):
    """A plan whose segments never cross the given column runs and whose work items are ordered by
    the runs' arrival group (isplib_b200_plan_build_grouped): the plan of the fused gather + SpMM
    kernel.  Usable with every forward entry point."""

    def __init__(self, rowptr32: torch.Tensor, col32: torch.Tensor, run_start, run_group, n_groups: int, seg_len: int = 0):
        assert rowptr32.is_cuda and rowptr32.dtype == torch.int32 and rowptr32.is_contiguous()
        assert col32.dtype == torch.int32 and col32.is_contiguous()
        self.m = rowptr32.numel() - 1
        self.nnz = int(col32.numel())
        n_runs = len(run_start)
        assert n_runs == len(run_group) and n_runs >= 1
        nbytes = ctypes.c_size_t(0)
        check(lib().isplib_b200_plan_grouped_bytes(self.m, self.nnz, seg_len, n_runs, ctypes.byref(nbytes)), "plan_grouped_bytes")
        self.buf = _dev_bytes(nbytes.value, rowptr32.device)
        self.info = PlanInfo()
        rs = (ctypes.c_int32 * n_runs)(*[int(v) for v in run_start])
        rg = (ctypes.c_int32 * n_runs)(*[int(v) for v in run_group])
        ends = (ctypes.c_int64 * int(n_groups))()
        check(lib().isplib_b200_plan_build_grouped(self.m, self.nnz, _p(rowptr32), _p(col32), seg_len, n_runs, rs, rg,
                                                   int(n_groups), _aligned_ptr(self.buf), nbytes.value,
                                                   ctypes.byref(self.info), ends, _stream(rowptr32.device)),
              "plan_build_grouped")
        self.n_groups = int(n_groups)
        self.group_item_end = [int(v) for v in ends]


def spmm_csr_gather(reduce, rowptr32, col32, val, x_gathered, plan, *, world: int, rank: int,
                    peer_x, peer_arrive, peer_credit, owner_group, my_group_at_peer, slice_rows: int, status,
                    epoch: int, parity_launch: int, tile_mode: bool = False, phase: int = 0,
                    copy_ctas: int = 0, variant: int = VARIANT_AUTO, out=None, arg_out=None, row_divisor=None,
                    edge_ids=None, arg_sentinel: Optional[int] = None, bias=None, addend=None,
                    addend_scale: float = 1.0, relu: bool = False, spmm_flags: int = 0, arg_col=None, arg_val=None):
    """Fused all-gather + SpMM (isplib_b200_spmm_csr_gather).  x_gathered: the LOCAL [world *
    slice_rows, K] buffer of this step's parity whose own slice already holds this step's rows;
    peer_x / peer_arrive / peer_credit: the device addresses (ints) of every rank's buffer / arrival
    counters / credit words as mapped into this process."""
    code = REDUCE_CODE[reduce] if isinstance(reduce, str) else int(reduce)
    assert x_gathered.is_cuda and x_gathered.dtype == torch.float32 and x_gathered.dim() == 2 and x_gathered.stride(1) == 1
    M, nnz = plan.m, plan.nnz
    N, K = x_gathered.shape
    is_arg = code in (MAX, MIN)
    if out is None:
        out = torch.empty((M, K), dtype=torch.float32, device=x_gathered.device)
    if is_arg and arg_out is None:
        arg_out = torch.empty((M, K), dtype=torch.int64, device=x_gathered.device)
    ws_bytes = ctypes.c_size_t(0)
    check(lib().isplib_b200_spmm_workspace_bytes(ctypes.byref(plan.info), K, code, ctypes.byref(ws_bytes)),
          "spmm_workspace_bytes")
    ws = _dev_bytes(ws_bytes.value, x_gathered.device)
    gd = GatherDesc()
    n_groups = 1 if tile_mode else plan.n_groups
    gd.world, gd.rank, gd.n_groups, gd.copy_ctas = world, rank, n_groups, copy_ctas
    gd.tile_mode, gd.phase = (1 if tile_mode else 0), int(phase)
    arrs = [(ctypes.c_void_p * world)(*[int(v) for v in a]) for a in (peer_x, peer_arrive, peer_credit)]
    gd.peer_x, gd.peer_arrive, gd.peer_credit = arrs
    og = (ctypes.c_int32 * world)(*[int(v) for v in owner_group])
    mg = (ctypes.c_int32 * world)(*[int(v) for v in my_group_at_peer])
    gie = (ctypes.c_int64 * n_groups)(*(plan.group_item_end if not tile_mode else [plan.info.num_items]))
    gd.owner_group, gd.my_group_at_peer, gd.group_item_end = og, mg, gie
    gd.slice_rows = slice_rows
    gd.status, gd.epoch = status.data_ptr(), int(epoch) & 0xFFFFFFFF
    gd.parity_launch = int(parity_launch) & 0xFFFFFFFF
    epi = Epilogue()
    epi.bias = None if bias is None else bias.data_ptr()
    if addend is not None:
        epi.addend = addend.data_ptr()
        epi.ld_addend = addend.stride(0) if addend.size(0) > 1 else max(K, addend.stride(0))
        epi.addend_scale = float(addend_scale)
    if arg_col is not None:
        epi.arg_col = arg_col.data_ptr()
        if arg_val is not None:
            epi.arg_val = arg_val.data_ptr()
    ldx = x_gathered.stride(0) if N > 1 else max(K, x_gathered.stride(0))
    ldo = out.stride(0) if M > 1 else max(K, out.stride(0))
    st = lib().isplib_b200_spmm_csr_gather(
        code, M, N, K, nnz, _p(rowptr32), _p(col32), _p(val), _p(x_gathered), ldx, _p(out), ldo,
        _p(arg_out) if is_arg else None, ctypes.byref(plan.info), plan.ptr, _aligned_ptr(ws), ws_bytes.value,
        variant, spmm_flags | (FLAG_RELU if relu else 0), _p(row_divisor), _p(edge_ids),
        nnz if arg_sentinel is None else int(arg_sentinel), ctypes.byref(epi), ctypes.byref(gd), _stream(x_gathered.device))
    check(st, "spmm_csr_gather")
    return out, (arg_out if is_arg else None)


def variant_names():
    L = lib()
    return [L.isplib_b200_variant_name(v).decode() for v in range(L.isplib_b200_variant_count())]


def spmm_csr(reduce, rowptr32, col32, val, x, plan: Plan, variant: int = VARIANT_AUTO, *,
             out=None, arg_out=None, flags: int = 0, row_divisor=None, edge_ids=None,
             arg_sentinel: Optional[int] = None, bias=None, addend=None, addend_scale: float = 1.0,
             relu: bool = False, arg_col=None, arg_val=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """out[, arg_out] = REDUCE(A, x) through isplib_b200_spmm_csr(_ex / _fused).

    bias [K], addend [M,K] (x addend_scale), relu: the fused caller epilogue; arg_col / arg_val
    ([M,K] int32 / fp32, contiguous like out): the auxiliary max/min outputs for the backward."""
    code = REDUCE_CODE[reduce] if isinstance(reduce, str) else int(reduce)
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    M, nnz = plan.m, plan.nnz
    N, K = x.shape
    is_arg = code in (MAX, MIN)
    if out is None:
        out = torch.empty((M, K), dtype=torch.float32, device=x.device)
    if is_arg and arg_out is None:
        # arg_out shares out's row stride (include/isplib_b200.h: "row stride ldo")
        ld = out.stride(0) if M > 1 else K
        arg_out = torch.empty((M, max(ld, K)), dtype=torch.int64, device=x.device)[:, :K]
    ws_bytes = ctypes.c_size_t(0)
    check(lib().isplib_b200_spmm_workspace_bytes(ctypes.byref(plan.info), K, code, ctypes.byref(ws_bytes)),
          "spmm_workspace_bytes")
    ws = _dev_bytes(ws_bytes.value, x.device)
    ldx = x.stride(0) if N > 1 else max(K, x.stride(0))
    ldo = out.stride(0) if M > 1 else max(K, out.stride(0))
    common = (code, M, N, K, nnz, _p(rowptr32), _p(col32), _p(val), _p(x), ldx, _p(out), ldo,
              _p(arg_out) if is_arg else None, ctypes.byref(plan.info), plan.ptr, _aligned_ptr(ws), ws_bytes.value)
    if bias is not None or addend is not None or relu or arg_col is not None:
        epi = Epilogue()
        epi.bias = None if bias is None else bias.data_ptr()
        if addend is not None:
            assert addend.is_cuda and addend.dtype == torch.float32 and addend.stride(1) == 1
            epi.addend = addend.data_ptr()
            epi.ld_addend = addend.stride(0) if addend.size(0) > 1 else max(K, addend.stride(0))
            epi.addend_scale = float(addend_scale)
        if arg_col is not None:
            assert arg_col.dtype == torch.int32 and (M <= 1 or arg_col.stride(0) == ldo)
            epi.arg_col = arg_col.data_ptr()
            if arg_val is not None:
                assert arg_val.dtype == torch.float32 and (M <= 1 or arg_val.stride(0) == ldo)
                epi.arg_val = arg_val.data_ptr()
        st = lib().isplib_b200_spmm_csr_fused(*common, variant, flags | (FLAG_RELU if relu else 0), _p(row_divisor),
                                              _p(edge_ids), nnz if arg_sentinel is None else int(arg_sentinel),
                                              ctypes.byref(epi), _stream(x.device))
    elif flags or row_divisor is not None or edge_ids is not None or arg_sentinel is not None:
        st = lib().isplib_b200_spmm_csr_ex(*common, variant, flags, _p(row_divisor), _p(edge_ids),
                                           nnz if arg_sentinel is None else int(arg_sentinel), _stream(x.device))
    else:
        st = lib().isplib_b200_spmm_csr(*common, variant, _stream(x.device))
    check(st, "spmm_csr")
    return out, (arg_out if is_arg else None)


def spmm_autotune(reduce, rowptr32, col32, val, x, plan: Plan, iters: int = 3):
    """(best_variant, [ms per variant]) -- isplib_b200_spmm_autotune."""
    code = REDUCE_CODE[reduce] if isinstance(reduce, str) else int(reduce)
    M, nnz = plan.m, plan.nnz
    N, K = x.shape
    is_arg = code in (MAX, MIN)
    out = torch.empty((M, K), dtype=torch.float32, device=x.device)
    arg_out = torch.empty((M, K), dtype=torch.int64, device=x.device) if is_arg else None
    ws_bytes = ctypes.c_size_t(0)
    check(lib().isplib_b200_spmm_workspace_bytes(ctypes.byref(plan.info), K, code, ctypes.byref(ws_bytes)),
          "spmm_workspace_bytes")
    ws = _dev_bytes(ws_bytes.value, x.device)
    nv = lib().isplib_b200_variant_count()
    times = (ctypes.c_float * nv)()
    best = ctypes.c_int(0)
    check(lib().isplib_b200_spmm_autotune(code, M, N, K, nnz, _p(rowptr32), _p(col32), _p(val), _p(x), x.stride(0),
                                          _p(out), K, _p(arg_out), ctypes.byref(plan.info), plan.ptr,
                                          _aligned_ptr(ws), ws_bytes.value, iters, ctypes.byref(best), times,
                                          _stream(x.device)), "spmm_autotune")
    return best.value, list(times)


def csr_transpose(rowptr32, col32, n: int):
    """(colptr, row_t, csr2csc), all int32 on the device -- isplib_b200_csr_transpose."""
    m, nnz = rowptr32.numel() - 1, col32.numel()
    dev = rowptr32.device
    colptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    row_t = torch.empty(nnz, dtype=torch.int32, device=dev)
    csr2csc = torch.empty(nnz, dtype=torch.int32, device=dev)
    ws_bytes = ctypes.c_size_t(0)
    check(lib().isplib_b200_csr_transpose_workspace_bytes(m, n, nnz, ctypes.byref(ws_bytes)), "transpose_ws")
    ws = _dev_bytes(ws_bytes.value, dev)
    check(lib().isplib_b200_csr_transpose(m, n, nnz, _p(rowptr32), _p(col32), _p(colptr), _p(row_t), _p(csr2csc),
                                          _aligned_ptr(ws), ws_bytes.value, _stream(dev)), "csr_transpose")
    return colptr, row_t, csr2csc


def permute_values(val, csr2csc, row_t, rowptr32, mean_weights: bool):
    nnz = csr2csc.numel()
    out = torch.empty(nnz, dtype=torch.float32, device=csr2csc.device)
    check(lib().isplib_b200_permute_values(nnz, _p(val), _p(csr2csc), _p(row_t), _p(rowptr32),
                                           1 if mean_weights else 0, _p(out), _stream(csr2csc.device)),
          "permute_values")
    return out


def spmm_arg_backward(col32, val, x, arg, grad_out, n: int, need_grad_x=True, need_grad_val=False,
                      arg_sentinel: Optional[int] = None):
    M, K = grad_out.shape
    nnz = col32.numel()
    dev = grad_out.device
    gx = torch.empty((n, K), dtype=torch.float32, device=dev) if need_grad_x else None
    gv = torch.empty(nnz, dtype=torch.float32, device=dev) if need_grad_val else None
    check(lib().isplib_b200_spmm_arg_backward(M, n, K, nnz, _p(col32), _p(val), _p(x), K if x is None else x.stride(0),
                                              _p(arg), arg.stride(0), nnz if arg_sentinel is None else arg_sentinel,
                                              _p(grad_out), grad_out.stride(0), _p(gx), K, _p(gv), 1, _stream(dev)),
          "spmm_arg_backward")
    return gx, gv


def spmm_arg_backward_aux(arg_col, arg_val, grad_out, n: int, binned: bool = False):
    """grad_x from the forward's auxiliary outputs -- isplib_b200_spmm_arg_backward_aux, or
    (binned=True) the partition-then-apply variant for a grad_x far beyond L2."""
    M, K = grad_out.shape
    dev = grad_out.device
    gx = torch.empty((n, K), dtype=torch.float32, device=dev)
    if binned:
        nb = ctypes.c_size_t(0)
        check(lib().isplib_b200_spmm_arg_backward_binned_workspace_bytes(M, n, K, ctypes.byref(nb)), "binned_ws")
        ws = _dev_bytes(nb.value, dev)
        check(lib().isplib_b200_spmm_arg_backward_binned(M, n, K, _p(arg_col), _p(arg_val), arg_col.stride(0) if M > 1 else K,
                                                         _p(grad_out), grad_out.stride(0) if M > 1 else K, _p(gx), K, 1,
                                                         _aligned_ptr(ws), nb.value, _stream(dev)), "spmm_arg_backward_binned")
        return gx
    check(lib().isplib_b200_spmm_arg_backward_aux(M, n, K, _p(arg_col), _p(arg_val), arg_col.stride(0) if M > 1 else K,
                                                  _p(grad_out), grad_out.stride(0) if M > 1 else K, _p(gx), K, 1,
                                                  _stream(dev)), "spmm_arg_backward_aux")
    return gx


def narrow_i64_to_i32(src: torch.Tensor) -> torch.Tensor:
    assert src.is_cuda and src.dtype == torch.int64 and src.is_contiguous()
    dst = torch.empty(src.shape, dtype=torch.int32, device=src.device)
    flag = torch.zeros(1, dtype=torch.int32, device=src.device)
    check(lib().isplib_b200_narrow_i64_to_i32(src.numel(), _p(src), _p(dst), _p(flag), _stream(src.device)),
          "narrow_i64_to_i32")
    if int(flag.item()) != 0:
        raise IsplibError(256, "narrow_i64_to_i32 (value out of int32 range)")
    return dst


def sddmm_csr(rowptr32, col32, a, x, plan: Plan, mean_scale: bool = False) -> torch.Tensor:
    """out_val[e] = <a[row(e)], x[col[e]]> (/ max(deg,1)) -- isplib_b200_sddmm_csr."""
    assert a.is_cuda and x.is_cuda and a.dtype == torch.float32 and x.dtype == torch.float32
    assert a.stride(1) == 1 and x.stride(1) == 1 and a.shape[1] == x.shape[1]
    M, K = a.shape
    N = x.shape[0]
    out = torch.empty(plan.nnz, dtype=torch.float32, device=x.device)
    check(lib().isplib_b200_sddmm_csr(M, N, K, plan.nnz, _p(rowptr32), _p(col32), _p(a), max(a.stride(0), K), _p(x),
                                      max(x.stride(0), K), 1 if mean_scale else 0, _p(out), ctypes.byref(plan.info),
                                      plan.ptr, _stream(x.device)), "sddmm_csr")
    return out


def coo_to_csr(row32, col32, val, m: int, n: int):
    """(rowptr, col, val, perm) int32/fp32 on the device, stable by (row, col) -- isplib_b200_coo_to_csr."""
    nnz = row32.numel()
    dev = row32.device
    rowptr = torch.empty(m + 1, dtype=torch.int32, device=dev)
    col_out = torch.empty(nnz, dtype=torch.int32, device=dev)
    perm = torch.empty(nnz, dtype=torch.int32, device=dev)
    val_out = None if val is None else torch.empty(nnz, dtype=torch.float32, device=dev)
    ws_bytes = ctypes.c_size_t(0)
    check(lib().isplib_b200_coo_to_csr_workspace_bytes(m, n, nnz, ctypes.byref(ws_bytes)), "coo_to_csr_ws")
    ws = _dev_bytes(ws_bytes.value, dev)
    check(lib().isplib_b200_coo_to_csr(m, n, nnz, _p(row32), _p(col32), _p(val), _p(rowptr), _p(col_out), _p(val_out),
                                       _p(perm), _aligned_ptr(ws), ws_bytes.value, _stream(dev)), "coo_to_csr")
    return rowptr, col_out, val_out, perm

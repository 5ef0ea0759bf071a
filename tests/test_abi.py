"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/isplib_b200.h declares; argument validation that needs no device."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "isplib_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isplib_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from isplib_b200 import capi
    lib = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/isplib_b200.h but not exported"
    assert sorted(capi.EXPORTS) == syms, "capi.EXPORTS out of sync with the header"
    assert lib.isplib_b200_abi_version() == 2


def test_no_torch_types_in_the_c_abi():
    text = open(os.path.join(ROOT, "include", "isplib_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "at::" not in code and "#include <cuda" not in code


def test_status_strings_and_variant_table():
    from isplib_b200 import capi
    lib = capi.lib()
    assert lib.isplib_b200_status_string(0) == b"success"
    assert b"workspace" in lib.isplib_b200_status_string(-1)
    names = capi.variant_names()
    assert len(names) == lib.isplib_b200_variant_count() >= 4 and len(set(names)) == len(names)
    assert lib.isplib_b200_variant_name(10_000) == b"invalid"


def test_host_side_argument_validation():
    """These paths return before touching the device."""
    from isplib_b200 import capi
    lib = capi.lib()
    n = ctypes.c_size_t(0)
    assert lib.isplib_b200_plan_bytes(-1, 0, 0, ctypes.byref(n)) == 256
    assert lib.isplib_b200_plan_bytes(10, 2**31, 0, ctypes.byref(n)) == 256          # nnz must fit int32
    assert lib.isplib_b200_plan_bytes(1000, 50_000, 0, ctypes.byref(n)) == 0 and n.value > 1000 * 4 * 3
    info = capi.PlanInfo(m=4, nnz=10, seg_len=256, num_items=4, num_split_items=3)
    assert lib.isplib_b200_spmm_workspace_bytes(ctypes.byref(info), 128, capi.SUM, ctypes.byref(n)) == 0
    sum_bytes = n.value
    assert lib.isplib_b200_spmm_workspace_bytes(ctypes.byref(info), 128, capi.MAX, ctypes.byref(n)) == 0
    assert n.value > sum_bytes >= 3 * 128 * 4
    assert lib.isplib_b200_spmm_workspace_bytes(ctypes.byref(info), 128, 9, ctypes.byref(n)) == 128
    # unknown FusedMM message -> FUSEDMM_NO_OPT_IMPL (csrc/fusedMM.h:114)
    assert lib.isplib_b200_fusedmm_csr_host(0x11101, 1, 1, 1, 1.0, 0, 1, 1, None, None, None, None, None, 1,
                                            None, 1, 0.0, None, 1, None) == 128


def test_plan_info_struct_layout_matches_header():
    from isplib_b200 import capi
    assert ctypes.sizeof(capi.PlanInfo) == 8 + 8 + 4 + 4 + 8 * 5 + 8


def test_variant_ids_quoted_in_profiles_and_bench_stay_valid():
    """profiles/README.md and bench.py quote variant ids / names; new variants must be appended
    to the table, never inserted, and the ncu traffic table may only name existing variants."""
    import importlib.util
    import os
    from isplib_b200 import capi, synth
    names = capi.variant_names()
    assert names[7] == "seg/w4/u4/kt64" and names[22] == "lean256/w4/kt64" and names[24] == "lean256/w4/kt64/seq"
    assert len(set(names)) == len(names)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for (shape, k, reduce, variant), traffic in bench.NCU_TRAFFIC_BYTES.items():
        assert variant in names and traffic > 0 and reduce in ("sum", "mean", "max", "min") and k > 0
        assert shape in synth.SHAPES


def test_variant_default_follows_the_measured_rules():
    """Shape-only default (no GPU needed, pointers are only checked for alignment):
    sum/mean -> lean 256-bit kernel, 64-wide sequential slabs once X exceeds L2; rows that are
    only 16-byte aligned -> lean128; max/min -> the same lean kernels from K = 64 up, seg below."""
    from isplib_b200 import capi
    L = capi.lib()
    names = capi.variant_names()
    ptr = 1 << 20                                    # 32-byte aligned fake address
    SUM, MAX = capi.REDUCE_CODE["sum"], capi.REDUCE_CODE["max"]

    def default(reduce, n, k, ldx=None, x=ptr):
        return names[L.isplib_b200_variant_default(reduce, n, k, ldx or k, k, x, ptr, 100.0)]

    assert default(SUM, 232965, 64) == "lean256/w4/kfull"
    assert default(SUM, 232965, 128) == "lean256/w4/kt64/seq"
    assert default(SUM, 1569960, 200) == "lean256/w4/kfull"          # ragged tile, HBM regime
    assert default(SUM, 2449029, 100) == "lean128/w4/kfull"          # rows only 16-byte aligned
    assert default(SUM, 2449029, 47, ldx=48) == "lean256/w4/kfull"   # padded odd width
    assert default(SUM, 1000, 64, x=ptr + 16) == "lean128/w4/kfull"  # operand only 16-byte aligned
    assert default(MAX, 232965, 128) == "lean256/w4/kt64"            # since r2: lean body at 32 warps/SM
    assert default(MAX, 232965, 256) == "lean256/w4/kt64/seq"
    assert default(MAX, 232965, 32).startswith("seg/")               # narrow rows stay with seg/*
    assert default(MAX, 1569960, 200) == "lean256/w4/kfull"
    assert default(SUM, 1000, 7).startswith("seg/")                  # scalar rows

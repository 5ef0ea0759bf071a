"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/isplib_b200.h declares; argument validation that needs no device."""
import ctypes
import os
import re

import pytest

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
    assert lib.isplib_b200_abi_version() == 1


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

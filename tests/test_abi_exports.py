"""The C-ABI library loads without a GPU and exports every function include/rkstiff_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rkstiff_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(rks_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("rks_plan_create", "rks_plan_destroy", "rks_update_coeffs", "rks_stage", "rks_nl",
                 "rks_error_control", "rks_read_ctrl", "rks_run_trials", "rks_run_fixed", "rks_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from rkstiff_b200 import _build
    _build.build_library()
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/rkstiff_b200.h but not exported"
    lib.rks_abi_version.restype = ctypes.c_int
    assert lib.rks_abi_version() == 1


def test_geometry_queries_need_no_gpu():
    from rkstiff_b200 import _abi
    lib = _abi.lib
    stages = {"IF4": 4, "ETD4": 4, "ETD5": 6, "IF34": 4, "ETD34": 4, "ETD35": 6, "IF45DP": 6}
    nbuf = {"IF4": 4, "ETD4": 4, "ETD5": 6, "IF34": 5, "ETD34": 5, "ETD35": 6, "IF45DP": 7}
    for m, mid in _abi.METHOD_IDS.items():
        assert lib.rks_num_stages(mid) == stages[m]
        assert lib.rks_num_nl_buffers(mid) == nbuf[m]
        assert lib.rks_is_adaptive(mid) == int(m in ("IF34", "ETD34", "ETD35", "IF45DP"))
        small = lib.rks_workspace_bytes(mid, 4, 513, 513, 0)
        big = lib.rks_workspace_bytes(mid, 8, 513, 513, 0)
        assert 0 < small < big
    assert lib.rks_workspace_bytes(99, 4, 513, 513, 0) == 0
    assert lib.rks_workspace_bytes(1, 4, 513, 100, 0) == 0          # lin_op must have n_c or batch*n_c entries


def test_no_cpu_fallback_without_cuda():
    """The product path fails loudly when there is no CUDA device (no oracle / CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import rkstiff_b200 as rk
    lin = torch.zeros(16, dtype=torch.float64)
    with pytest.raises(ValueError):
        rk.ETD4(lin, lambda v: v)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rkstiff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} refers to oracle/"

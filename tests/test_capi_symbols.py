"""The C-ABI library loads without a GPU and exports every symbol include/ecw_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ecw_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ecw_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_header(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    import ecw_cc_b200
    dll = ctypes.CDLL(ecw_cc_b200.LIB_PATH)
    for s in syms:
        assert hasattr(dll, s), "missing export %s" % s
    assert b"sm_100a" in built_lib.ecw_version()


def test_no_gpu_fails_loudly(built_lib):
    """Without a device every compute entry point must fail (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    assert built_lib.ecw_ctx_create(ctypes.byref(h), 3, 4) == 0
    rc = built_lib.ecw_ccsd_gamma(h, None, None, None, None, None, None)
    assert rc != 0
    assert b"no CUDA device" in built_lib.ecw_last_error(h)
    built_lib.ecw_ctx_destroy(h)
    import ecw_cc_b200
    import pytest
    with pytest.raises(ecw_cc_b200.EcwError):
        ecw_cc_b200.DeviceEris.synthetic(3, 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ecw_cc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f

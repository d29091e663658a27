import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """The product library, compiled in-tree if it is not there yet."""
    import ecw_cc_b200
    ecw_cc_b200.build()
    return ecw_cc_b200.lib


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["int8", "int8_all", "dmma"])
def engine(request, monkeypatch):
    """GEMM engine of the contraction plans: "int8" = product default (large GEMMs and the packed-vvvv
    ladders on the INT8 tcgen05 pipe), "int8_all" = every unbatched GEMM forced onto it, "dmma" = FP64
    DMMA kernels only.  Parity tests run under all three."""
    monkeypatch.setenv("ECW_GEMM", "dmma" if request.param == "dmma" else "int8")
    monkeypatch.setenv("ECW_INT8_MIN_FLOPS", "-1" if request.param == "int8_all" else "2e10")
    return request.param

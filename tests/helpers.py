"""Shared test helpers (CPU side)."""
import ctypes
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = [("upd", None, False), ("eq", None, True), ("l1upd", 1e-3, False), ("l1eq", 1e-3, True)]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def plan_json(lib, o, v, func, flags, rank=0, world=1, int8_digits=0, min_flops=-1.0, vvvv_planes=False,
              ovvv_planes=False, splitk_min_k=None, cut_cache_min=None):
    """Plan of one call.  int8_digits > 0: INT8 tensor-core engine for every unbatched GEMM with
    2MNK >= min_flops; vvvv_planes: the packed vvvv is bound as digit planes (needs a device for the
    real entry point, so the flag is flipped through the test hook ecw_ctx_test_assume_vvvv_planes)."""
    h = ctypes.c_void_p()
    assert lib.ecw_ctx_create(ctypes.byref(h), o, v) == 0
    if world > 1:
        assert lib.ecw_ctx_set_shard(h, rank, world) == 0
    if int8_digits:
        assert lib.ecw_ctx_set_gemm(h, int8_digits, float(min_flops)) == 0
        if splitk_min_k is not None:
            assert lib.ecw_ctx_set_int8_splitk(h, splitk_min_k) == 0
        if vvvv_planes:
            assert lib.ecw_ctx_test_assume_vvvv_planes(h) == 0
        if ovvv_planes:
            assert lib.ecw_ctx_test_assume_ovvv_planes(h) == 0
        if cut_cache_min is not None:
            assert lib.ecw_ctx_test_cut_cache_min(h, cut_cache_min) == 0
    try:
        n = 1 << 24
        buf = ctypes.create_string_buffer(n)
        r = lib.ecw_plan_dump(h, func.encode(), flags, buf, n)
        assert r > 0, lib.ecw_last_error(h)
        return json.loads(buf.value.decode())
    finally:
        lib.ecw_ctx_destroy(h)


def eris_slots(er):
    """numpy versions of the constant device layouts (oracle/refactored_np.DeviceErisSpec)."""
    from oracle import refactored_np as R
    E = R.DeviceErisSpec(er)
    return dict(oooo=E.oooo.copy(), ooov=E.ooov.copy(), oovv=E.oovv.copy(), oovv_ph=E.oovv_ph,
                ovov_ph=E.ovov_ph, ovvv=E.ovvv.copy(), oooo_p=E.oooo_p, oovv_p=E.oovv_p,
                ovvv_p=E.ovvv_p, vvvv_p=E.vvvv_p)


def flags_of(alpha, equation, antisym=True):
    return (1 if alpha is not None else 0) | (2 if equation else 0) | (4 if antisym else 0)

"""Randomised check of the contraction engine's lowering (Plan::contract, csrc/plan.cpp) on the CPU: random binary
einsum specs, extents and memory layouts, lowered under the DMMA engine, the INT8 engine (every unbatched GEMM on the
digit-plane route) and the INT8 engine with split-K forced on; each plan is replayed by tests/plan_interp.py and
compared with numpy.einsum.  Covers what ecw_op_contract (the CCS class, the GCC intermediate getters) can be asked."""
import ctypes
import json

import numpy as np
import pytest

from plan_interp import Interp


class _T(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("nd", ctypes.c_int32), ("dim", ctypes.c_int64 * 6), ("str", ctypes.c_int64 * 6)]


def _desc(shape, strides):
    t = _T()
    t.ptr = None
    t.nd = len(shape)
    for i, (d, s) in enumerate(zip(shape, strides)):
        t.dim[i], t.str[i] = d, s
    return t


def _layout(rng, shape):
    """random storage order (+ occasional padding of the slowest axis): element strides and the flat buffer size"""
    order = list(rng.permutation(len(shape)))
    strides = [0] * len(shape)
    s = 1
    for ax in order:
        strides[ax] = s
        s *= shape[ax] + (1 if rng.random() < 0.2 else 0)
    return strides, max(s, 1)


def _view(buf, shape, strides):
    return np.lib.stride_tricks.as_strided(buf, shape=shape, strides=[8 * s for s in strides])


def _plan(lib, h, alpha, A, sa, B, sb, beta, C, sc):
    n = 1 << 22
    buf = ctypes.create_string_buffer(n)
    r = lib.ecw_plan_dump_contract(h, alpha, ctypes.byref(A), sa.encode(), ctypes.byref(B), sb.encode(), beta,
                                   ctypes.byref(C), sc.encode(), buf, n)
    assert r > 0, lib.ecw_last_error(h)
    return json.loads(buf.value.decode())


@pytest.mark.parametrize("mode", ["dmma", "int8", "int8_splitk"])
def test_random_contractions(built_lib, mode):
    lib = built_lib
    rng = np.random.default_rng({"dmma": 1, "int8": 2, "int8_splitk": 3}[mode])
    h = ctypes.c_void_p()
    assert lib.ecw_ctx_create(ctypes.byref(h), 4, 6) == 0
    if mode != "dmma":
        assert lib.ecw_ctx_set_gemm(h, 6, -1.0) == 0
        assert lib.ecw_ctx_set_int8_splitk(h, 64 if mode == "int8_splitk" else 1 << 40) == 0
    kinds = {"gemm": 0, "oz_gemm": 0, "split-K": 0}
    labels = "abcdefgh"
    for trial in range(150):
        ni, nj, nk = rng.integers(0, 3), rng.integers(0, 3), rng.integers(0, 3 if mode != "int8_splitk" else 2) + (mode == "int8_splitk")
        if ni + nk == 0 or nj + nk == 0 or ni + nj == 0:
            continue
        lab = list(rng.permutation(list(labels))[: ni + nj + nk])
        I, J, K = lab[:ni], lab[ni:ni + nj], lab[ni + nj:]
        big = mode == "int8_splitk"
        dims = {c: int(rng.integers(1, 7)) for c in lab}
        if big:
            dims[K[0]] = int(rng.choice([64, 96, 128]))            # a long contraction index: split-K kicks in
        sa = "".join(rng.permutation(I + K))
        sb = "".join(rng.permutation(J + K))
        sc = "".join(rng.permutation(I + J))
        shp = lambda s: tuple(dims[c] for c in s)
        (stA, nA), (stB, nB), (stC, nC) = _layout(rng, shp(sa)), _layout(rng, shp(sb)), _layout(rng, shp(sc))
        bufA, bufB, bufC = rng.standard_normal(nA), rng.standard_normal(nB), rng.standard_normal(nC)
        alpha, beta = float(rng.choice([1.0, -0.5, 2.0])), float(rng.choice([0.0, 1.0, -0.25]))
        A, B, C = _view(bufA, shp(sa), stA), _view(bufB, shp(sb), stB), _view(bufC, shp(sc), stC)
        ref = alpha * np.einsum("%s,%s->%s" % (sa, sb, sc), A, B) + beta * C
        pl = _plan(lib, h, alpha, _desc(shp(sa), stA), sa, _desc(shp(sb), stB), sb, beta, _desc(shp(sc), stC), sc)
        for op in pl["ops"]:
            if op["kind"] in kinds:
                kinds[op["kind"]] += 1
            if op["kind"] == "oz_gemm" and "split-K" in op["note"]:
                kinds["split-K"] += 1
        got_buf = bufC.copy()
        Interp(pl, {"a0": bufA, "a1": bufB, "b0": got_buf}).run()
        got = _view(got_buf, shp(sc), stC)
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(got - ref).max() < 1e-11 * scale, (mode, trial, sa, sb, sc, dims)
    lib.ecw_ctx_destroy(h)
    assert kinds["gemm"] + kinds["oz_gemm"] > 50
    if mode != "dmma":
        assert kinds["oz_gemm"] > 20
    if mode == "int8_splitk":
        assert kinds["split-K"] > 10

"""The sampled-element oracle (oracle/ccsd_columns.py) — the independent arithmetic the benchmark shape (40,400) is
checked against on the GPU (tests/test_gpu_bench_shape.py) — equals the full oracle (oracle/ccsd_np.py, pinned to the
unmodified reference) on every element it can produce, in every output mode; and the C generator of the synthetic
inputs (oracle/csrc/synth_c.c) is bit-identical to oracle/synth.py."""
import itertools

import numpy as np
import pytest

from helpers import MODES
from oracle import synth, synth_fast
from oracle.ccsd_columns import ColumnOracle
from oracle.ccsd_np import OracleGCC


def test_fast_generator_bit_exact():
    o, v = 5, 7
    er = synth.SynthEris(o, v)
    p = synth_fast.SynthProvider(o, v)
    assert synth_fast.lib() is not None, "gcc -fopenmp is part of the image"
    for name in ("oooo", "ooov", "oovv", "ovov"):
        assert np.array_equal(getattr(p, name), getattr(er, name)), name
    assert np.array_equal(p.ovvv_m(1, 4), er.ovvv[1:4])
    assert np.array_equal(p.ovvv_x1(3), er.ovvv[:, 3]) and np.array_equal(p.ovvv_x2(6), er.ovvv[:, :, 6])
    assert np.array_equal(p.ovvv_ef(2, 5), er.ovvv[:, :, 2, 5])
    assert np.array_equal(p.vvvv_ab(2, 5), er.vvvv[2, 5]) and np.array_equal(p.vvvv_x3(4), er.vvvv[:, :, 4])
    for x, y in zip(synth_fast.amplitudes(o, v), synth.amplitudes(o, v)):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("o,v,sym", [(4, 6, True), (5, 8, False)])
def test_columns_equal_full_oracle(o, v, sym):
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    if not sym:          # amplitudes without antisymmetry (what an L1-regularised iteration produces, Q11)
        rng = np.random.default_rng(7)
        t2 = t2 + 0.01 * rng.standard_normal(t2.shape)
        l2 = l2 + 0.01 * rng.standard_normal(l2.shape)
    fsp = synth.fsp(o, v)
    pairs = list(itertools.product(range(v), range(v)))            # every (a, b), a == b and a > b included
    full = OracleGCC(er)
    for provider in (synth_fast.ArrayProvider(er), synth_fast.SynthProvider(o, v)):
        col = ColumnOracle(provider, pairs)
        assert col.xs == list(range(v))
        for tag, alpha, eq in MODES:
            for f in (fsp, None):
                a1, a2 = full.tupdate(t1, t2, fsp=f, alpha=alpha, equation=eq)
                b1, b2 = col.tupdate(t1, t2, fsp=f, alpha=alpha, equation=eq)
                want = np.stack([a2[:, :, a, b] for a, b in pairs])
                assert np.abs(a1 - b1).max() < 1e-13 and np.abs(want - b2).max() < 1e-13, (tag, "T")
                a1, a2 = full.lupdate(t1, t2, l1, l2, fsp=f, alpha=alpha, equation=eq)
                b1, b2 = col.lupdate(t1, t2, l1, l2, fsp=f, alpha=alpha, equation=eq)
                want = np.stack([a2[:, :, a, b] for a, b in pairs])
                assert np.abs(a1 - b1).max() < 1e-13 and np.abs(want - b2).max() < 1e-13, (tag, "L")

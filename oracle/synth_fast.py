"""Fast provider of the function-defined synthetic inputs — TEST INFRASTRUCTURE ONLY.

Same definition as oracle/synth.py (and the CUDA generators, ecw_cc_b200/csrc/synth.cu), evaluated by a small C
library (oracle/csrc/synth_c.c, compiled here with gcc -fopenmp into oracle/_build/) so that slices of the benchmark
shape (40,400) — o v^3 = 2.6e9 elements of ovvv, v^3 rows of vvvv — cost seconds.  Falls back to the numpy generator
when no C compiler is around.  `SynthProvider` hands the column oracle (oracle/ccsd_columns.py) the blocks of
`Eris.geris` (Eris.py:132-150) on demand, never a dense vvvv.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import synth

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "synth_c.c")
_SO = os.path.join(_HERE, "_build", "libecw_oracle_synth.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -shared; returns the path or None when it cannot be built."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    try:
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _SO])
    except (OSError, subprocess.CalledProcessError):
        return None
    return _SO


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            _lib = False
        else:
            d = ctypes.CDLL(so)
            p64 = ctypes.POINTER(ctypes.c_int64)
            d.ecw_oracle_eri_block.argtypes = [ctypes.c_int64, ctypes.c_double, p64, ctypes.c_int64, p64, ctypes.c_int64,
                                               p64, ctypes.c_int64, p64, ctypes.c_int64, ctypes.c_void_p]
            d.ecw_oracle_eri_block.restype = None
            d.ecw_oracle_doubles.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double,
                                             ctypes.c_void_p]
            d.ecw_oracle_doubles.restype = None
            _lib = d
    return _lib or None


def eri_block(n, ps, qs, rs, ss, scale=synth.ERI_SCALE):
    d = lib()
    if d is None:
        return synth.eri_block(n, ps, qs, rs, ss, scale)
    idx = [np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in (ps, qs, rs, ss)]
    out = np.empty(tuple(len(x) for x in idx), dtype=np.float64)
    args = []
    for x in idx:
        args += [x.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(x)]
    d.ecw_oracle_eri_block(n, float(scale), *args, out.ctypes.data)
    return out


def doubles(nocc, nvir, seed, scale=0.02):
    d = lib()
    if d is None:
        return synth.doubles(nocc, nvir, seed, scale)
    out = np.empty((nocc, nocc, nvir, nvir), dtype=np.float64)
    d.ecw_oracle_doubles(nocc, nvir, seed, float(scale), out.ctypes.data)
    return out


def amplitudes(nocc, nvir):
    return (synth.singles(nocc, nvir, synth.SEED_T1), doubles(nocc, nvir, synth.SEED_T2),
            synth.singles(nocc, nvir, synth.SEED_L1), doubles(nocc, nvir, synth.SEED_L2))


class SynthProvider(object):
    """Integral blocks of the synthetic workload on demand (what oracle/ccsd_columns.ColumnOracle asks for)."""

    def __init__(self, nocc, nvir, scale=synth.ERI_SCALE, cache_ovvv=None):
        """cache_ovvv: keep ovvv (8 o v^3 bytes) in host memory once generated; None = when a third of the free RAM
        holds it."""
        self.nocc, self.nvir, self.scale = nocc, nvir, scale
        if cache_ovvv is None:
            try:
                import psutil
                cache_ovvv = 8.0 * nocc * nvir ** 3 < psutil.virtual_memory().available / 3.0
            except ImportError:
                cache_ovvv = 8.0 * nocc * nvir ** 3 < 4e9
        self._cache = {} if cache_ovvv else None
        n = self.n = nocc + nvir
        self.o = np.arange(nocc)
        self.v = np.arange(nocc, n)
        self.fock = synth.fock(nocc, nvir)
        o, v = self.o, self.v
        self.oooo = self.blk(o, o, o, o)
        self.ooov = self.blk(o, o, o, v)
        self.oovv = self.blk(o, o, v, v)
        self.ovov = self.blk(o, v, o, v)

    def blk(self, p, q, r, s):
        return eri_block(self.n, p, q, r, s, self.scale)

    def ovvv_m(self, m0, m1):
        """ovvv[m0:m1, :, :, :]"""
        if self._cache is None:
            return self.blk(self.o[m0:m1], self.v, self.v, self.v)
        for m in range(m0, m1):
            if m not in self._cache:
                self._cache[m] = self.blk(self.o[m:m + 1], self.v, self.v, self.v)[0]
        if m1 - m0 == 1:
            return self._cache[m0][None]
        return np.stack([self._cache[m] for m in range(m0, m1)])

    def ovvv_x1(self, a):
        """ovvv[:, a, :, :]  (o, v, v)"""
        return self.blk(self.o, self.v[a:a + 1], self.v, self.v)[:, 0]

    def ovvv_x2(self, a):
        """ovvv[:, :, a, :]  (o, v, v)"""
        return self.blk(self.o, self.v, self.v[a:a + 1], self.v)[:, :, 0]

    def ovvv_ef(self, e, f):
        """ovvv[:, :, e, f]  (o, v)"""
        return self.blk(self.o, self.v, self.v[e:e + 1], self.v[f:f + 1])[:, :, 0, 0]

    def vvvv_ab(self, a, b):
        """vvvv[a, b, :, :]  (v, v)"""
        return self.blk(self.v[a:a + 1], self.v[b:b + 1], self.v, self.v)[0, 0]

    def vvvv_x3(self, a):
        """vvvv[:, :, a, :]  (v, v, v)"""
        return self.blk(self.v, self.v, self.v[a:a + 1], self.v)[:, :, 0]


class ArrayProvider(object):
    """The same interface over a dense container with the attribute surface of `Eris.geris` (small sizes)."""

    def __init__(self, eris):
        self.nocc = eris.nocc
        self.fock = np.asarray(eris.fock)
        self.nvir = self.fock.shape[0] - self.nocc
        self.oooo, self.ooov, self.oovv, self.ovov = (np.asarray(getattr(eris, k)) for k in ("oooo", "ooov", "oovv", "ovov"))
        self._ovvv, self._vvvv = np.asarray(eris.ovvv), np.asarray(eris.vvvv)

    def ovvv_m(self, m0, m1):
        return self._ovvv[m0:m1]

    def ovvv_x1(self, a):
        return self._ovvv[:, a]

    def ovvv_x2(self, a):
        return self._ovvv[:, :, a]

    def ovvv_ef(self, e, f):
        return self._ovvv[:, :, e, f]

    def vvvv_ab(self, a, b):
        return self._vvvv[a, b]

    def vvvv_x3(self, a):
        return self._vvvv[:, :, a]


class FastSynthEris(object):
    """`synth.SynthEris` (attribute surface of `Eris.geris`, Eris.py:132-154) filled by the C generator: the canonical
    blocks the hot path reads plus the aliases the reference formulas name."""

    def __init__(self, nocc, nvir, scale=synth.ERI_SCALE):
        n = nocc + nvir
        o, v = np.arange(nocc), np.arange(nocc, n)
        blk = lambda a, b, c, d: eri_block(n, a, b, c, d, scale)  # noqa: E731
        self.nocc = nocc
        self.fock = synth.fock(nocc, nvir)
        self.oooo, self.ooov, self.oovv = blk(o, o, o, o), blk(o, o, o, v), blk(o, o, v, v)
        self.ovov, self.ovvo, self.ovvv = blk(o, v, o, v), blk(o, v, v, o), blk(o, v, v, v)
        self.vvvv = blk(v, v, v, v)
        self.vovv, self.oovo, self.voov = blk(v, o, v, v), blk(o, o, v, o), blk(v, o, o, v)
        self.mo_occ = np.concatenate([np.ones(nocc), np.zeros(nvir)])
        self.EHF = -1.0
        self.orbspin = np.arange(n) % 2

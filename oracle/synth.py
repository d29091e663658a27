"""Function-defined synthetic inputs (numpy side) — TEST INFRASTRUCTURE ONLY.

Every element is a pure function of its indices (splitmix64 hash), so the CPU
oracle, the CUDA generator kernels (`ecw_cc_b200/csrc/synth.cu`) and every
shard of a multi-GPU run produce bit-identical values without a dense `vvvv`
ever existing.  The tensors carry the symmetries of the reference integral
container (`Eris.py:128`: <pq||rs> = -<qp||rs> = -<pq||sr> = <rs||pq>) and the
block names of `Eris.py:132-150`.

Definition (shared with the CUDA side, see DESIGN.md "Synthetic inputs"):
    z = splitmix64(key + seed * 0x9E3779B97F4A7C15)          (mod 2^64)
    u(seed, key) = (z >> 11) * 2^-52 - 1          in [-1, 1), exact in FP64
    <pq||rs> = 0 if p==q or r==s else
               sgn * ERI_SCALE * u(1, min(B,K) * n^2 + max(B,K)),
               B = p'n+q', K = r'n+s' after sorting each pair ascending
    fock  = diag(eps),  eps_i = -2 + 1.5 i/(o-1),  eps_a = 0.5 + 2.5 a/(v-1)
    V_pq  = 0.05 u(2, min(p,q) n + max(p,q));   fsp = fock - V
    t1,l1 = 0.05 u(3|4, i v + a)
    t2,l2 (i<j,a<b) = 0.02 u(5|6, ((i o + j) v + a) v + b), exactly 0 when the
               low 3 bits of z are 0, then antisymmetrised.
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
ERI_SCALE = 0.01
SEED_ERI, SEED_V, SEED_T1, SEED_L1, SEED_T2, SEED_L2 = 1, 2, 3, 4, 5, 6


def splitmix64(key, seed):
    """Vectorised splitmix64 finaliser; key: uint64 array, seed: int."""
    with np.errstate(over="ignore"):
        x = key.astype(np.uint64) + np.uint64(seed) * GOLDEN
        x = x + GOLDEN
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def unit(z):
    """Map hash bits to [-1, 1) exactly."""
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0


def eri_block(n, ps, qs, rs, ss, scale=ERI_SCALE):
    """Antisymmetrised <pq||rs> for index ranges (each a 1-D int array of
    absolute spin-orbital indices).  Returns array [len(ps),len(qs),len(rs),len(ss)]."""
    p = np.asarray(ps, dtype=np.int64)[:, None, None, None]
    q = np.asarray(qs, dtype=np.int64)[None, :, None, None]
    r = np.asarray(rs, dtype=np.int64)[None, None, :, None]
    s = np.asarray(ss, dtype=np.int64)[None, None, None, :]
    p, q, r, s = np.broadcast_arrays(p, q, r, s)
    sgn = np.where(p < q, 1.0, -1.0) * np.where(r < s, 1.0, -1.0)
    bra = np.minimum(p, q) * n + np.maximum(p, q)
    ket = np.minimum(r, s) * n + np.maximum(r, s)
    key = (np.minimum(bra, ket) * (n * n) + np.maximum(bra, ket)).astype(np.uint64)
    val = sgn * (scale * unit(splitmix64(key, SEED_ERI)))
    val = np.where((p == q) | (r == s), 0.0, val)
    return np.ascontiguousarray(val)


def orbital_energies(nocc, nvir):
    eo = -2.0 + 1.5 * np.arange(nocc) / max(nocc - 1, 1)
    ev = 0.5 + 2.5 * np.arange(nvir) / max(nvir - 1, 1)
    return np.concatenate([eo, ev])


def fock(nocc, nvir):
    return np.diag(orbital_energies(nocc, nvir))


def vexp(nocc, nvir, scale=0.05):
    n = nocc + nvir
    p = np.arange(n, dtype=np.int64)[:, None]
    q = np.arange(n, dtype=np.int64)[None, :]
    key = (np.minimum(p, q) * n + np.maximum(p, q)).astype(np.uint64)
    return scale * unit(splitmix64(key, SEED_V))


def fsp(nocc, nvir):
    """Dressed one-body operator fock - Vexp (cf. Solver_GS.py:692)."""
    return fock(nocc, nvir) - vexp(nocc, nvir)


def singles(nocc, nvir, seed, scale=0.05):
    i = np.arange(nocc, dtype=np.int64)[:, None]
    a = np.arange(nvir, dtype=np.int64)[None, :]
    return scale * unit(splitmix64((i * nvir + a).astype(np.uint64), seed))


def doubles(nocc, nvir, seed, scale=0.02):
    i = np.arange(nocc, dtype=np.int64)[:, None, None, None]
    j = np.arange(nocc, dtype=np.int64)[None, :, None, None]
    a = np.arange(nvir, dtype=np.int64)[None, None, :, None]
    b = np.arange(nvir, dtype=np.int64)[None, None, None, :]
    i, j, a, b = np.broadcast_arrays(i, j, a, b)
    lo_i, hi_i = np.minimum(i, j), np.maximum(i, j)
    lo_a, hi_a = np.minimum(a, b), np.maximum(a, b)
    key = (((lo_i * nocc + hi_i) * nvir + lo_a) * nvir + hi_a).astype(np.uint64)
    z = splitmix64(key, seed)
    val = scale * unit(z)
    val = np.where((z & np.uint64(7)) == np.uint64(0), 0.0, val)
    sgn = np.where(i < j, 1.0, -1.0) * np.where(a < b, 1.0, -1.0)
    val = np.where((i == j) | (a == b), 0.0, sgn * val)
    return np.ascontiguousarray(val)


class SynthEris:
    """Object with the attribute surface of `Eris.geris` (Eris.py:132-154)
    filled with the synthetic antisymmetrised integrals."""

    def __init__(self, nocc, nvir, with_vvvv=True, scale=ERI_SCALE):
        n = nocc + nvir
        o = np.arange(nocc)
        v = np.arange(nocc, n)
        blk = lambda a, b, c, d: eri_block(n, a, b, c, d, scale)  # noqa: E731
        self.nocc = nocc
        self.fock = fock(nocc, nvir)
        self.oooo = blk(o, o, o, o)
        self.ooov = blk(o, o, o, v)
        self.oovv = blk(o, o, v, v)
        self.ovov = blk(o, v, o, v)
        self.ovvo = blk(o, v, v, o)
        self.ovvv = blk(o, v, v, v)
        if with_vvvv:
            self.vvvv = blk(v, v, v, v)
        self.vooo = blk(v, o, o, o)
        self.vovo = blk(v, o, v, o)
        self.oovo = blk(o, o, v, o)
        self.vovv = blk(v, o, v, v)
        self.vvoo = blk(v, v, o, o)
        self.vvvo = blk(v, v, v, o)
        self.voov = blk(v, o, o, v)
        self.ovoo = blk(o, v, o, o)
        self.mo_occ = np.concatenate([np.ones(nocc), np.zeros(nvir)])
        self.EHF = -1.0
        self.orbspin = np.arange(n) % 2


def amplitudes(nocc, nvir):
    """(t1, t2, l1, l2) of the synthetic workload."""
    return (singles(nocc, nvir, SEED_T1), doubles(nocc, nvir, SEED_T2),
            singles(nocc, nvir, SEED_L1), doubles(nocc, nvir, SEED_L2))

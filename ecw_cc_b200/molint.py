"""Integral source without PySCF: Gaussian one- and two-electron integrals over s, p and (spherical) d
shells (McMurchie-Davidson), restricted Hartree-Fock, and the spin-orbital antisymmetrised MO integrals of the reference's
`Eris.geris` container (Eris.py:24-154: `<pq||rs> = <pq|rs> - <pq|sr>`, blocks `oooo ... vvvv`, `fock = diag(mo_energy)`),
built the way `Main.ECW.__init__` does (Main.py:151-217: RHF -> GHF spin orbitals in `[a, b, a, b, ...]` order).

Host-side numpy on purpose: this is the producer of the path's inputs (a 13-function basis for config 1, H2O/6-31G),
not the path.  Basis sets embedded for H, C, N, O: the 6-31G family — 6-31G, 6-31G*, 6-31G**, 6-31+G*, 6-31++G** ...
(Pople's standard polarisation / diffuse exponents; five spherical d functions per shell, PySCF's default) — and
cc-pVDZ for H, C, O.  Anchor: E_HF(H2O/6-31G) = -75.9839 (ECW_CC/__init__.py:39).
"""
import numpy as np
from scipy.special import hyp1f1

ANGSTROM = 1.0 / 0.52917721092                 # bohr per Angstrom (the value PySCF uses)

# 6-31G: per element a list of (kind, exponents, s coefficients, p coefficients)
BASIS_631G = {
    1: [("S", [18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733], None),
        ("S", [0.1612778], [1.0], None)],
    6: [("S", [3047.5249, 457.36952, 103.94869, 29.210155, 9.2866630, 3.1639270],
         [0.0018347, 0.0140373, 0.0688426, 0.2321844, 0.4679413, 0.3623120], None),
        ("SP", [7.8682724, 1.8812885, 0.5442493], [-0.1193324, -0.1608542, 1.1434564], [0.0689991, 0.3164240, 0.7443083]),
        ("SP", [0.1687144], [1.0], [1.0])],
    7: [("S", [4173.5110, 627.45790, 142.90210, 40.234330, 12.820210, 4.3904370],
         [0.0018348, 0.0139950, 0.0685870, 0.2322410, 0.4690700, 0.3604550], None),
        ("SP", [11.626358, 2.7162800, 0.7722180], [-0.1149610, -0.1691180, 1.1458520], [0.0675800, 0.3239070, 0.7408950]),
        ("SP", [0.2120313], [1.0], [1.0])],
    8: [("S", [5484.6717, 825.23495, 188.04696, 52.964500, 16.897570, 5.7996353],
         [0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.4701930, 0.3585209], None),
        ("SP", [15.539616, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.1307670], [0.0708743, 0.3397528, 0.7271586]),
        ("SP", [0.2700058], [1.0], [1.0])],
}


# polarisation / diffuse functions of the 6-31G family: d exponent ('*'), diffuse sp exponent ('+') of the heavy atoms,
# p exponent ('**') and diffuse s exponent ('++') of hydrogen
POPLE_D = {6: 0.8, 7: 0.8, 8: 0.8}
POPLE_DIFFUSE_SP = {6: 0.0438, 7: 0.0639, 8: 0.0845}
POPLE_H_P, POPLE_H_DIFFUSE_S = 1.1, 0.036

# cc-pVDZ (Dunning 1989): general contractions written out per contracted function
BASIS_CCPVDZ = {
    1: [("S", [13.01, 1.962, 0.4446], [0.019685, 0.137977, 0.478148], None),
        ("S", [0.122], [1.0], None),
        ("P", [0.727], None, [1.0])],
    6: [("S", [6665.0, 1000.0, 228.0, 64.71, 21.06, 7.495, 2.797, 0.5215],
         [0.000692, 0.005329, 0.027077, 0.101718, 0.27474, 0.448564, 0.285074, 0.015204], None),
        ("S", [6665.0, 1000.0, 228.0, 64.71, 21.06, 7.495, 2.797, 0.5215],
         [-0.000146, -0.001154, -0.005725, -0.023312, -0.063955, -0.149981, -0.127262, 0.544529], None),
        ("S", [0.1596], [1.0], None),
        ("P", [9.439, 2.002, 0.5456], None, [0.038109, 0.20948, 0.508557]),
        ("P", [0.1517], None, [1.0]),
        ("D", [0.55], None, [1.0])],
    8: [("S", [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013],
         [0.00071, 0.00547, 0.027837, 0.1048, 0.283062, 0.448719, 0.270952, 0.015458], None),
        ("S", [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013],
         [-0.00016, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368], None),
        ("S", [0.3023], [1.0], None),
        ("P", [17.7, 3.854, 1.046], None, [0.043018, 0.228913, 0.508728]),
        ("P", [0.2753], None, [1.0]),
        ("D", [1.185], None, [1.0])],
}

# real solid harmonics of a d shell over Cartesian monomials, in PySCF's order (xy, yz, z2, xz, x2-y2); factors are
# relative to the norm of an xy-type Gaussian, N = (2a/pi)^(3/4) * 4a
_S3 = 1.0 / (2.0 * np.sqrt(3.0))
D_SPHERICAL = [[((1, 1, 0), 1.0)], [((0, 1, 1), 1.0)], [((0, 0, 2), 2 * _S3), ((2, 0, 0), -_S3), ((0, 2, 0), -_S3)],
               [((1, 0, 1), 1.0)], [((2, 0, 0), 0.5), ((0, 2, 0), -0.5)]]


def basis_shells(name, z):
    """Shell list [(kind, exponents, s coefficients, p-or-d coefficients)] of element z in the named basis."""
    key = name.lower().replace("-", "").replace("(d)", "*").replace("(d,p)", "**")
    if key == "ccpvdz":
        if z not in BASIS_CCPVDZ:
            raise NotImplementedError("cc-pVDZ is embedded for H, C, O")
        return BASIS_CCPVDZ[z]
    import re
    m = re.fullmatch(r"631(\+{0,2})g(\*{0,2})", key)
    if m is None:
        raise NotImplementedError("embedded basis sets: 6-31G family (6-31G, 6-31G*, 6-31+G*, 6-31++G**, ...) and cc-pVDZ")
    plus, star = len(m.group(1)), len(m.group(2))
    shells = list(BASIS_631G[z])
    if z == 1:
        if plus == 2:
            shells.append(("S", [POPLE_H_DIFFUSE_S], [1.0], None))
        if star == 2:
            shells.append(("P", [POPLE_H_P], None, [1.0]))
    else:
        if plus >= 1:
            shells.append(("SP", [POPLE_DIFFUSE_SP[z]], [1.0], [1.0]))
        if star >= 1:
            shells.append(("D", [POPLE_D[z]], None, [1.0]))
    return shells


class Molecule(object):
    """atoms: [(Z or element symbol, (x, y, z)), ...] in Angstrom (Main.py:104-109 style); basis: '6-31g', '6-31+g*',
    '6-31++g**', 'cc-pvdz', ... (see `basis_shells`)."""
    SYMBOLS = {"H": 1, "C": 6, "N": 7, "O": 8}

    def __init__(self, atoms, basis="6-31g", charge=0):
        for a in atoms:
            if (a[0].capitalize() if isinstance(a[0], str) else int(a[0])) not in (set(self.SYMBOLS) | set(self.SYMBOLS.values())):
                raise NotImplementedError("element %r: basis sets are embedded for H, C, N, O" % (a[0],))
        self.Z = np.array([self.SYMBOLS[a[0].capitalize()] if isinstance(a[0], str) else a[0] for a in atoms], dtype=np.float64)
        self.R = np.array([a[1] for a in atoms], dtype=np.float64) * ANGSTROM
        self.nelec = int(self.Z.sum()) - charge
        if self.nelec % 2:
            raise NotImplementedError("closed shells only (RHF)")
        # primitive Cartesian Gaussians: centre, exponent, (lx,ly,lz), coefficient incl. normalisation, AO index
        cen, ex, lmn, coef, ao = [], [], [], [], []
        nao = 0
        for z, r in zip(self.Z, self.R):
            for kind, exps, cs, cp in basis_shells(basis, int(z)):
                parts = {"S": ((0, cs),), "P": ((1, cp),), "SP": ((0, cs), (1, cp)), "D": ((2, cp),)}[kind]
                for ang, cc in parts:
                    if ang == 2:                               # five spherical d functions
                        comps = D_SPHERICAL
                    else:
                        comps = [[((0, 0, 0), 1.0)]] if ang == 0 else [[(l3, 1.0)] for l3 in ((1, 0, 0), (0, 1, 0), (0, 0, 1))]
                    for comp in comps:
                        for a, c in zip(exps, cc):
                            norm = (2 * a / np.pi) ** 0.75 * (2 * np.sqrt(a)) ** ang      # s / p / xy-type d norm
                            for l3, f in comp:
                                cen.append(r); ex.append(a); lmn.append(l3); coef.append(c * norm * f); ao.append(nao)
                        nao += 1
        self.cen, self.ex = np.array(cen), np.array(ex)
        self.lmn, self.coef, self.ao = np.array(lmn), np.array(coef), np.array(ao)
        self.nao = nao
        self.lmax = int(self.lmn.sum(1).max())
        # Tables of general contractions (cc-pVDZ) list contracted functions that are not unit normalised: rescale
        # those.  Segmented Pople functions are normalised to ~1e-7 as published and are left exactly as tabulated.
        sii = self._self_overlaps()
        scale = np.where(np.abs(sii - 1.0) > 1e-4, 1.0 / np.sqrt(sii), 1.0)
        self.coef = self.coef * scale[self.ao]
        d = self.R[:, None, :] - self.R[None, :, :]
        rr = np.sqrt((d ** 2).sum(-1))
        iu = np.triu_indices(len(self.Z), 1)
        self.e_nuc = float((self.Z[:, None] * self.Z[None, :])[iu].dot(1.0 / rr[iu]))
        self._origin = np.zeros(3)
        self._ints = None


    def _self_overlaps(self):
        """<mu|mu> of every AO from its own primitives (all on one centre: products of 1-D Gaussian moments)."""
        def dfact(n):                                             # (n-1)!! for even n >= 0
            return float(np.prod(np.arange(n - 1, 0, -2))) if n > 0 else 1.0
        out = np.zeros(self.nao)
        for mu in range(self.nao):
            k = np.where(self.ao == mu)[0]
            for i in k:
                for j in k:
                    n3 = self.lmn[i] + self.lmn[j]
                    if (n3 % 2).any():
                        continue
                    pq = self.ex[i] + self.ex[j]
                    val = (np.pi / pq) ** 1.5
                    for n in n3:
                        val *= dfact(n) / (2 * pq) ** (n // 2)
                    out[mu] += self.coef[i] * self.coef[j] * val
        return out

    # -- the slice of PySCF's `gto.Mole` surface that exp_pot.Exp / utilities touch (exp_pot.py:90-108) ---------------
    def atom_charges(self):
        return self.Z.copy()

    def atom_coords(self):
        return self.R.copy()                                      # Bohr, like PySCF

    def with_common_orig(self, origin):
        mol = self

        class _Origin(object):
            def __enter__(self_inner):
                self_inner.old = mol._origin
                mol._origin = np.asarray(origin, dtype=np.float64)
                return mol

            def __exit__(self_inner, *exc):
                mol._origin = self_inner.old
                return False
        return _Origin()

    def intor_symmetric(self, name, comp=None):
        if self._ints is None:
            self._ints = integrals(self)
        if name == "int1e_kin":
            return self._ints[1]
        if name == "int1e_nuc":
            return self._ints[2]
        if name == "int1e_ovlp":
            return self._ints[0]
        if name == "int1e_r":
            return dipole_integrals(self, self._origin)
        raise NotImplementedError("intor %r" % (name,))

    intor = intor_symmetric


def _boys(nmax, x):
    """F_n(x), n = 0..nmax, for an array x."""
    return np.stack([hyp1f1(n + 0.5, n + 1.5, -x) / (2 * n + 1) for n in range(nmax + 1)])


def _hermite_E(imax, jmax, p, XPA, XPB, mu, XAB):
    """1-D Hermite expansion coefficients E[i][j][t] (arrays over pairs), i <= imax, j <= jmax."""
    E = [[None] * (jmax + 1) for _ in range(imax + 1)]
    tmax = imax + jmax
    z = np.zeros_like(p)
    E[0][0] = [np.exp(-mu * XAB ** 2)] + [z] * (tmax + 1)
    for i in range(imax + 1):
        for j in range(jmax + 1):
            if i == 0 and j == 0:
                continue
            if i > 0:
                src, X = E[i - 1][j], XPA
            else:
                src, X = E[i][j - 1], XPB
            E[i][j] = [(src[t - 1] / (2 * p) if t > 0 else 0) + X * src[t] + (t + 1) * src[t + 1]
                       for t in range(tmax + 1)] + [z]
    return E


def _hermite_R(L, alpha, D):
    """Hermite Coulomb integrals R[t][u][v] (t+u+v <= L) for exponent alpha and distance vectors D[..., 3]."""
    r2 = (D ** 2).sum(-1)
    F = _boys(L, alpha * r2)
    Rn = {(0, 0, 0, n): (-2 * alpha) ** n * F[n] for n in range(L + 1)}
    X, Y, Z = D[..., 0], D[..., 1], D[..., 2]

    def get(t, u, v, n):
        key = (t, u, v, n)
        if key in Rn:
            return Rn[key]
        if t > 0:
            val = X * get(t - 1, u, v, n + 1) + ((t - 1) * get(t - 2, u, v, n + 1) if t > 1 else 0)
        elif u > 0:
            val = Y * get(t, u - 1, v, n + 1) + ((u - 1) * get(t, u - 2, v, n + 1) if u > 1 else 0)
        else:
            val = Z * get(t, u, v - 1, n + 1) + ((v - 1) * get(t, u, v - 2, n + 1) if v > 1 else 0)
        Rn[key] = val
        return val
    return {(t, u, v): get(t, u, v, 0) for t in range(L + 1) for u in range(L + 1 - t) for v in range(L + 1 - t - u)}


def _pairs(mol):
    """Unique primitive pairs I >= J and the matrix that adds a pair quantity into the AO pairs (mu nu) and (nu mu)."""
    I, J = np.tril_indices(len(mol.ex))
    nao = mol.nao
    M = np.zeros((len(I), nao * nao))
    k = np.arange(len(I))
    np.add.at(M, (k, mol.ao[I] * nao + mol.ao[J]), 1.0)
    off = I != J
    np.add.at(M, (k[off], mol.ao[J[off]] * nao + mol.ao[I[off]]), 1.0)
    return I, J, M


def integrals(mol):
    """Overlap, kinetic, nuclear-attraction [nao, nao] and electron-repulsion (mu nu|la si) [nao]*4, AO basis.
    McMurchie-Davidson over the unique primitive pairs (I >= J) and the lower triangle of pair-pairs."""
    I, J, M = _pairs(mol)
    a, b = mol.ex[I], mol.ex[J]
    p = a + b
    mu = a * b / p
    A, B = mol.cen[I], mol.cen[J]
    P = (a[:, None] * A + b[:, None] * B) / p[:, None]
    la, lb = mol.lmn[I], mol.lmn[J]
    cc = mol.coef[I] * mol.coef[J]
    lm = mol.lmax
    nt = 2 * lm + 1                                               # Hermite orders 0..2*lmax per direction
    # 1-D coefficients with j up to l_b + 2 (kinetic energy)
    Ed = [_hermite_E(lm, lm + 2, p, P[:, d] - A[:, d], P[:, d] - B[:, d], mu, A[:, d] - B[:, d]) for d in range(3)]

    def pick(d, dj, t):
        """E^{la_d, lb_d + dj}_t per pair (lb_d + dj may be -1 or -2: zero)."""
        out = np.zeros(len(p))
        for i in range(lm + 1):
            for j in range(0, lm + 3):
                if t > i + j:
                    continue
                m = (la[:, d] == i) & (lb[:, d] + dj == j)
                if m.any():
                    out[m] = Ed[d][i][j][t][m]
        return out
    S1 = [pick(d, 0, 0) * np.sqrt(np.pi / p) for d in range(3)]
    T1 = []
    for d in range(3):
        j = lb[:, d]
        T1.append((-2 * b ** 2 * pick(d, 2, 0) + b * (2 * j + 1) * pick(d, 0, 0) - 0.5 * j * (j - 1) * pick(d, -2, 0))
                  * np.sqrt(np.pi / p))
    Sp = S1[0] * S1[1] * S1[2]
    Tp = T1[0] * S1[1] * S1[2] + S1[0] * T1[1] * S1[2] + S1[0] * S1[1] * T1[2]
    # Hermite tensor of every pair: Eab[pair, t, u, v], t + u + v <= 2*lmax
    Eab = np.zeros((len(p), nt, nt, nt))
    e1 = [[pick(d, 0, t) for t in range(nt)] for d in range(3)]
    for t in range(nt):
        for u in range(nt):
            for v in range(nt):
                if t + u + v <= 2 * lm:
                    Eab[:, t, u, v] = e1[0][t] * e1[1][u] * e1[2][v]
    Vp = np.zeros(len(p))
    for Zc, C in zip(mol.Z, mol.R):
        R = _hermite_R(2 * lm, p, P - C)
        acc = np.zeros(len(p))
        for (t, u, v), r in R.items():
            acc += Eab[:, t, u, v] * r
        Vp -= Zc * 2 * np.pi / p * acc
    nao = mol.nao
    S, T, V = ((M.T @ (cc * x)).reshape(nao, nao) for x in (Sp, Tp, Vp))
    # electron repulsion: (ab|cd) = 2 pi^2.5 / (p q sqrt(p+q)) sum E_ab(tuv) (-1)^(t'+u'+v') E_cd(t'u'v') R(t+t',u+u',v+v')
    sign = np.array([[[(-1.0) ** (t + u + v) for v in range(nt)] for u in range(nt)] for t in range(nt)])
    Ecd = Eab * sign[None]
    idx = [(t, u, v) for t in range(nt) for u in range(nt) for v in range(nt) if t + u + v <= 2 * lm]
    W = np.zeros((len(p), len(p)))                                # pair-pair integrals, lower block triangle
    chunk = 64
    for s0 in range(0, len(p), chunk):
        s1 = min(len(p), s0 + chunk)
        sl, kt = slice(s0, s1), slice(0, s1)
        pq = p[sl, None] + p[None, kt]
        alpha = p[sl, None] * p[None, kt] / pq
        R = _hermite_R(4 * lm, alpha, P[sl, None, :] - P[None, kt, :])
        acc = np.zeros_like(alpha)
        for (t, u, v) in idx:
            ea = Eab[sl, t, u, v]
            if not ea.any():
                continue
            for (t2, u2, v2) in idx:
                ec = Ecd[kt, t2, u2, v2]
                if not ec.any():
                    continue
                acc += ea[:, None] * ec[None, :] * R[(t + t2, u + u2, v + v2)]
        W[sl, kt] = 2 * np.pi ** 2.5 / (p[sl, None] * p[None, kt] * np.sqrt(pq)) * acc * cc[sl, None] * cc[None, kt]
    W = np.tril(W) + np.tril(W, -1).T
    eri = M.T @ W @ M
    return S, T, V, eri.reshape(nao, nao, nao, nao)


def dipole_integrals(mol, origin=(0., 0., 0.)):
    """<mu| r - origin |nu>, [3, nao, nao] (the 'int1e_r' integrals the reference takes from PySCF with
    `mol.with_common_orig`, exp_pot.py:90-98): 1-D first moments (E_1 + (P - C) E_0) sqrt(pi/p) times the overlaps of
    the other two directions."""
    I, J, M = _pairs(mol)
    a, b = mol.ex[I], mol.ex[J]
    p = a + b
    mu = a * b / p
    A, B = mol.cen[I], mol.cen[J]
    P = (a[:, None] * A + b[:, None] * B) / p[:, None]
    la, lb = mol.lmn[I], mol.lmn[J]
    cc = mol.coef[I] * mol.coef[J]
    origin = np.asarray(origin, dtype=np.float64)
    S1, M1 = [], []
    for d in range(3):
        lm = mol.lmax
        E = _hermite_E(lm, lm, p, P[:, d] - A[:, d], P[:, d] - B[:, d], mu, A[:, d] - B[:, d])
        e0, e1 = np.zeros(len(p)), np.zeros(len(p))
        for i in range(lm + 1):
            for j in range(lm + 1):
                m = (la[:, d] == i) & (lb[:, d] == j)
                e0[m], e1[m] = E[i][j][0][m], E[i][j][1][m]
        S1.append(e0 * np.sqrt(np.pi / p))
        M1.append((e1 + (P[:, d] - origin[d]) * e0) * np.sqrt(np.pi / p))
    nao = mol.nao
    out = []
    for d in range(3):
        f = [M1[k] if k == d else S1[k] for k in range(3)]
        out.append((M.T @ (cc * f[0] * f[1] * f[2])).reshape(nao, nao))
    return np.stack(out)


def rhf(mol, ints=None, conv=1e-11, maxiter=100):
    """Closed-shell SCF with DIIS.  Returns (E_HF, mo_energy, mo_coeff, ints)."""
    S, T, V, eri = ints if ints is not None else integrals(mol)
    h = T + V
    nocc = mol.nelec // 2
    s, U = np.linalg.eigh(S)
    X = U / np.sqrt(s)
    e, C = np.linalg.eigh(X.T @ h @ X)
    C = X @ C
    D = 2 * C[:, :nocc] @ C[:, :nocc].T
    fs, es = [], []
    E_old = 0.0
    for it in range(maxiter):
        F = h + np.einsum("mnls,ls->mn", eri, D) - 0.5 * np.einsum("mlns,ls->mn", eri, D)
        E = 0.5 * np.sum(D * (h + F)) + mol.e_nuc
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        fs.append(F); es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:
            n = len(fs)
            Bm = -np.ones((n + 1, n + 1)); Bm[n, n] = 0.0
            for i in range(n):
                for j in range(n):
                    Bm[i, j] = np.sum(es[i] * es[j])
            rhs = np.zeros(n + 1); rhs[n] = -1.0
            try:
                c = np.linalg.solve(Bm, rhs)[:n]
                F = sum(ci * fi for ci, fi in zip(c, fs))
            except np.linalg.LinAlgError:
                pass
        e, C = np.linalg.eigh(X.T @ F @ X)
        C = X @ C
        D = 2 * C[:, :nocc] @ C[:, :nocc].T
        if abs(E - E_old) < conv and np.abs(err).max() < 1e-8:
            break
        E_old = E
    F = h + np.einsum("mnls,ls->mn", eri, D) - 0.5 * np.einsum("mlns,ls->mn", eri, D)
    E = 0.5 * np.sum(D * (h + F)) + mol.e_nuc
    e, C = np.linalg.eigh(X.T @ F @ X)
    return float(E), e, X @ C, (S, T, V, eri)


class geris(object):
    """The attribute surface of the reference's `Eris.geris` (Eris.py:132-154) from an RHF solution:
    spin orbitals in [alpha, beta, alpha, beta, ...] order (orbspin = 0,1,0,1,...), `fock = diag(mo_energy)`,
    `<pq||rs> = (pr|qs) - (ps|qr)` with the spin selection rules, all sixteen o/v blocks."""

    def __init__(self, mol, scf_result=None):
        ehf, e, C, (S, T, V, eri) = scf_result if scf_result is not None else rhf(mol)
        nmo = C.shape[1]
        mo = np.einsum("mp,mnls->pnls", C, eri)
        mo = np.einsum("nq,pnls->pqls", C, mo)
        mo = np.einsum("lr,pqls->pqrs", C, mo)
        mo = np.einsum("st,pqrs->pqrt", C, mo)                  # (pq|rs), spatial MOs, chemists' notation
        n = 2 * nmo
        sp = np.arange(n) // 2                                   # spatial index of spin orbital
        spin = np.arange(n) % 2
        g = mo[np.ix_(sp, sp, sp, sp)] * (spin[:, None, None, None] == spin[None, :, None, None]) \
            * (spin[None, None, :, None] == spin[None, None, None, :])       # (pq|rs) spin orbitals
        phys = g.transpose(0, 2, 1, 3)                           # <pq|rs> = (pr|qs)
        anti = phys - phys.transpose(0, 1, 3, 2)
        o = mol.nelec
        self.nocc = o
        self.fock = np.diag(np.repeat(e, 2))
        self.mo_energy = np.repeat(e, 2)
        self.mo_occ = np.concatenate([np.ones(o), np.zeros(n - o)])
        self.orbspin = spin
        self.mo_coeff = C
        nao = C.shape[0]
        self.mo_coeff_g = np.zeros((2 * nao, n))                 # G format: [[C, 0], [0, C]] with interleaved columns,
        self.mo_coeff_g[:nao, 0::2] = C                           # what scf.addons.convert_to_ghf(mf).mo_coeff holds
        self.mo_coeff_g[nao:, 1::2] = C
        self.EHF = self.e_hf = ehf
        sl = {"o": slice(0, o), "v": slice(o, n)}
        for name in ("oooo", "ooov", "oovv", "ovov", "ovvo", "ovvv", "vvvv", "vooo", "vovo", "oovo", "vovv", "vvoo", "vvvo",
                     "voov", "ovoo"):
            setattr(self, name, np.ascontiguousarray(anti[sl[name[0]], sl[name[1]], sl[name[2]], sl[name[3]]]))

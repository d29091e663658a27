"""The PySCF-free integral source (ecw_cc_b200/molint.py): anchors and structure of what it hands to the
path, and config 1 (H2O/6-31G) through the oracle against runs of the unmodified reference solver."""
import numpy as np
import pytest

from helpers import load_golden
from oracle.ccsd_np import OracleGCC
from oracle.make_golden_h2o import CASES, H2O
from oracle.make_golden_solver import target_rdm1
from oracle.solver_np import ExpMat, scf_loop


@pytest.fixture(scope="module")
def water():
    from ecw_cc_b200 import molint
    mol = molint.Molecule(H2O, "6-31g")
    scf = molint.rhf(mol)
    return molint, mol, scf


def test_rhf_anchor_and_integral_symmetries(water):
    molint, mol, (ehf, e, C, (S, T, V, eri)) = water
    assert mol.nao == 13 and mol.nelec == 10
    assert abs(ehf - (-75.9839)) < 5e-5                 # ECW_CC/__init__.py:39 (EHF = -7.59839e+01)
    assert abs(ehf - (-75.98394849810545)) < 1e-8       # value of this implementation (PySCF: -75.983948...)
    assert np.abs(np.diag(S) - 1).max() < 1e-6 and np.abs(S - S.T).max() < 1e-14
    assert np.abs(T - T.T).max() < 1e-12 and np.abs(V - V.T).max() < 1e-12
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)):
        assert np.abs(eri - eri.transpose(perm)).max() < 1e-12
    assert np.abs(C.T @ S @ C - np.eye(13)).max() < 1e-10
    assert e[0] < -20.5 and e[4] < 0 < e[5]             # O 1s core, 5 occupied orbitals


def test_geris_surface(water):
    molint, mol, scf = water
    er = molint.geris(mol, scf)
    o, n = er.nocc, er.fock.shape[0]
    assert (o, n - o) == (10, 16) and np.array_equal(er.fock, np.diag(np.diagonal(er.fock)))
    assert np.abs(er.oovv + er.oovv.transpose(1, 0, 2, 3)).max() < 1e-12
    assert np.abs(er.oovv + er.oovv.transpose(0, 1, 3, 2)).max() < 1e-12
    assert np.abs(er.vvvv - er.vvvv.transpose(2, 3, 0, 1)).max() < 1e-12
    assert np.abs(er.ovvo + er.ovov.transpose(0, 1, 3, 2)).max() < 1e-12          # ovvo[iabj] = -ovov[iajb]
    assert np.abs(er.voov + er.ovov.transpose(1, 0, 2, 3)).max() < 1e-12          # voov[akic] = -ovov[kaic]
    # Brillouin: the HF determinant's singles vanish, the MP2 energy is negative and of the known size
    e = np.diagonal(er.fock)
    d = e[:o, None, None, None] + e[None, :o, None, None] - e[None, None, o:, None] - e[None, None, None, o:]
    emp2 = 0.25 * np.sum(er.oovv ** 2 / d)
    assert -0.14 < emp2 < -0.12


def test_config1_oracle_reproduces_reference_runs(water):
    """H2O/6-31G ECW-CCSD ground states: oracle loop vs the unmodified reference solver (tests/golden/h2o_631g.npz)."""
    molint, mol, _ = water
    g = load_golden("h2o_631g.npz")
    ints = molint.integrals(mol)
    er = molint.geris(mol, (float(g["EHF"]), g["mo_energy"], g["mo_coeff"], ints))
    assert abs(float(g["L0_Ep"][-1]) - (-0.1353978855)) < 1e-9      # CCSD correlation energy of water in 6-31G
    tag, L, alpha, maxiter = CASES[1]
    out = scf_loop(OracleGCC(er), ExpMat(L, target_rdm1(10, 16)), L, alpha=alpha, conv_thres=float(g["conv_thres"]),
                   maxiter=maxiter)
    assert out[0] == str(g[tag + "_text"])
    assert np.abs(out[1] - g[tag + "_Ep"]).max() < 1e-11 and np.abs(out[4] - g[tag + "_rdm1"]).max() < 1e-10


def test_acetylene_rhf_is_reproducible():
    """C2H2/6-31G (config 2 in the 6-31G basis): the RHF solution stored with the sweep fixture is what molint gives."""
    from oracle.make_golden_c2h2 import acetylene
    g = load_golden("c2h2_631g_sweep.npz")
    mol, er, scf = acetylene()
    assert mol.nao == 22 and (er.nocc, er.fock.shape[0]) == (14, 44)
    assert abs(scf[0] - float(g["EHF"])) < 1e-9 and abs(scf[0] - (-76.79224)) < 1e-4
    assert np.abs(scf[1] - g["mo_energy"]).max() < 1e-7


def test_d_shells_and_basis_families():
    """Five spherical d functions per shell (PySCF's default): function counts of the named basis sets, rotational
    invariance of the RHF energy (mixes the d components), energies against known values, normalised AOs."""
    from ecw_cc_b200 import molint
    counts = {"6-31g": 13, "6-31g*": 18, "6-31+g*": 22, "6-31++g**": 30, "cc-pvdz": 24}
    for basis, nao in counts.items():
        assert molint.Molecule(H2O, basis).nao == nao, basis
    with pytest.raises(NotImplementedError):
        molint.Molecule(H2O, "def2-svp")
    rng = np.random.default_rng(0)
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    turned = [(z, tuple(Q @ np.array(r))) for z, r in H2O]
    # RHF/cc-pVDZ water at this geometry: -76.0268 (literature); 6-31+G*: value of this implementation
    for basis, want, tol in (("cc-pvdz", -76.0268, 1e-4), ("6-31+g*", -76.0162450889, 1e-8)):
        mol = molint.Molecule(H2O, basis)
        ints = molint.integrals(mol)
        e0 = molint.rhf(mol, ints)[0]
        assert abs(e0 - want) < tol, (basis, e0)
        assert np.abs(np.diag(ints[0]) - 1).max() < 1e-6
        assert abs(molint.rhf(molint.Molecule(turned, basis))[0] - e0) < 1e-10, basis
    er = molint.geris(molint.Molecule(H2O, "6-31+g*"))
    assert (er.nocc, er.fock.shape[0] - er.nocc) == (10, 34)          # the (10, 34) of SURVEY §8(f)-1


def test_basis_name_variants():
    from ecw_cc_b200 import molint
    kinds = lambda name, z: [sh[0] for sh in molint.basis_shells(name, z)]
    assert kinds("6-31G(d)", 8) == kinds("6-31g*", 8) == ["S", "SP", "SP", "D"]
    assert kinds("6-31+G(d,p)", 1) == ["S", "S", "P"] and kinds("6-31+G(d,p)", 6) == ["S", "SP", "SP", "SP", "D"]
    assert kinds("6-31++G**", 1) == ["S", "S", "S", "P"]
    assert kinds("cc-pVDZ", 6) == ["S", "S", "S", "P", "P", "D"]
    with pytest.raises(NotImplementedError):
        molint.basis_shells("cc-pvdz", 7)
    with pytest.raises(NotImplementedError):
        molint.Molecule([("Fe", (0., 0., 0.))])


def test_acetylene_ccpvdz_shape_of_config2():
    """C2H2/cc-pVDZ (config 2's named basis; carbon's cc-pVDZ table and hydrogen's p shell): 38 functions,
    (nocc, nvir) = (14, 62) as SURVEY §8(f)-1 quotes, RHF energy -76.8255 (literature, 5e-4) — one minute of integrals."""
    from ecw_cc_b200 import molint
    from oracle.make_golden_c2h2 import C2H2
    mol = molint.Molecule(C2H2, "cc-pvdz")
    ints = molint.integrals(mol)
    ehf, e, C, _ = molint.rhf(mol, ints)
    assert mol.nao == 38 and np.abs(np.diag(ints[0]) - 1).max() < 1e-6
    assert abs(ehf - (-76.82553727888)) < 1e-8 and abs(ehf - (-76.8255)) < 5e-4
    assert 2 * mol.nao - mol.nelec == 62

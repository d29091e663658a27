"""Golden vectors of the UNMODIFIED reference `exp_pot.Exp` (exp_pot.py:11-490) for every target kind the product's
mirror provides — 'mat' (GS / ES), 'Ek', 'v1e', 'dip', 'DEk', 'trdip' — on H2O/6-31G, the AO integrals coming from
`ecw_cc_b200.molint.Molecule` through the PySCF-like surface the reference touches (intor_symmetric, with_common_orig,
atom_charges, atom_coords).  Build container only:

    python -m oracle.make_golden_exp

tests/golden/exp_pot_h2o.npz: per call (state index pair) the potential, Delta, vmax and the calculated properties.
Inputs are function defined (`setup` below), so the fixture carries outputs only (plus the RHF orbitals).
"""
import os

import numpy as np

from . import ref_loader
from .make_golden_h2o import H2O

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CALLS = [(0, 0), (1, 1), (0, 1), (1, 0), (2, 2), (0, 2), (2, 0)]
WEIGHTS = [[0.1, 0.1, 0.03], [0.04, 0.07, 0.06, 0.02], [0.05, 0.01, 0.2]]


def setup(mol, mo_coeff_g, utilities):
    """exp_data for GS + 2 ES and the 'calculated' matrices handed to Vexp_update, all from seeded random numbers;
    `utilities` = the module whose Ekin/dipole/v1e turn the target matrices into target properties."""
    n = mo_coeff_g.shape[0]
    rng = np.random.default_rng(77)

    def sym(scale):
        a = scale * rng.standard_normal((n, n))
        return 0.5 * (a + a.T)
    occ = np.diag(np.concatenate([np.ones(10), np.zeros(n - 10)]))
    g_exp = occ + sym(0.02)
    g_es1, g_es2 = occ + sym(0.05), occ + sym(0.05)
    trl1, trr1 = 0.1 * rng.standard_normal((n, n)), 0.1 * rng.standard_normal((n, n))
    trl2, trr2 = 0.1 * rng.standard_normal((n, n)), 0.1 * rng.standard_normal((n, n))
    kw = dict(aobasis=False, mo_coeff=mo_coeff_g)
    ek_gs, ek1, ek2 = (float(utilities.Ekin(mol, g, **kw)) for g in (g_exp, g_es1, g_es2))
    trdip1 = list(utilities.dipole(mol, trr1, **kw) * utilities.dipole(mol, trl1, **kw))
    trdip2 = list(utilities.dipole(mol, trr2, **kw) * utilities.dipole(mol, trl2, **kw))
    exp_data = [[['mat', g_exp], ['DEk1', abs(ek1 - ek_gs)], ['dip', list(utilities.dipole(mol, g_exp, **kw))]],
                [['DEk1', ek1 - ek_gs], ['trdip', trdip1], ['v1e', float(utilities.v1e(mol, g_es1, **kw))], ['mat', g_es1]],
                [['Ek', ek2], ['trdip', trdip2], ['dip', list(utilities.dipole(mol, g_es2, **kw))]]]
    calc = {"g0": occ + sym(0.03), "g1": occ + sym(0.06), "g2": occ + sym(0.06),
            "l1": trl1 + 0.02 * rng.standard_normal((n, n)), "r1": trr1 + 0.02 * rng.standard_normal((n, n)),
            "l2": trl2 + 0.02 * rng.standard_normal((n, n)), "r2": trr2 + 0.02 * rng.standard_normal((n, n))}
    return exp_data, calc


def call_args(calc, index):
    """(rdm1, rdm1_add) of a Vexp_update call, as Solver_ES.py:262-287 passes them."""
    n, m = index
    if n == m:
        return calc["g%d" % n], calc["g0"]
    k = max(n, m)
    right, left = calc["r%d" % k], calc["l%d" % k]
    return (right, left) if m == 0 else (left, right)


def run(Exp, mol, mo_coeff_g, utilities):
    exp_data, calc = setup(mol, mo_coeff_g, utilities)
    out = {}
    for tag, L in (("w", WEIGHTS), ("s", 0.05)):
        vx = Exp(L if tag == "s" else [list(l) for l in WEIGHTS], exp_data, mol, mo_coeff_g, Ek_exp_GS=76.2, Ek_HF_GS=75.99)
        for (n, m) in CALLS:
            a, b = call_args(calc, (n, m))
            delta, vmax = vx.Vexp_update(a, b, (n, m))
            key = "%s_%d%d" % (tag, n, m)
            out[key + "_V"], out[key + "_Delta"], out[key + "_vmax"] = np.array(vx.Vexp[n, m]), delta, vmax
            out[key + "_V00"] = np.array(vx.Vexp[0, 0])
            out[key + "_prop"] = np.concatenate([np.atleast_1d(np.asarray(p[1], dtype=float)) for p in vx.prop_calc] or [np.zeros(0)])
        out[tag + "_EkGS"], out[tag + "_DEkGS"] = vx.Ek_calc_GS, vx.Delta_Ek_GS
    return out


def main():
    from ecw_cc_b200 import molint
    exp_pot, utilities = ref_loader.load("exp_pot", "utilities")
    g = np.load(os.path.join(OUT, "h2o_631g.npz"))
    mol = molint.Molecule(H2O, "6-31g")
    er = molint.geris(mol, (float(g["EHF"]), g["mo_energy"], g["mo_coeff"], molint.integrals(mol)))
    out = run(exp_pot.Exp, mol, er.mo_coeff_g, utilities)
    np.savez_compressed(os.path.join(OUT, "exp_pot_h2o.npz"), **out)
    for k in sorted(out):
        if k.endswith("_Delta") or k.endswith("_vmax"):
            print(k, out[k])


if __name__ == "__main__":
    main()

"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/ECW_CC, imported behind oracle/pyscf_stub) on the synthetic
inputs of oracle/synth.py.  Run in the build container only:

    python -m oracle.make_golden

The fixtures hold reference OUTPUTS (and the one non-function-defined input, a
random non-symmetric dressed Fock); inputs are regenerated from oracle/synth.py.
"""
import os

import numpy as np

from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
MODES = [("upd", None, False), ("eq", None, True), ("l1upd", 1e-3, False), ("l1eq", 1e-3, True)]


def ccsd_case(o, v):
    CCSD, RAW = ref_loader.load("CCSD", "CC_raw_equations")
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    rng = np.random.default_rng(1000 * o + v)
    fsp_ns = er.fock + 0.05 * rng.standard_normal(er.fock.shape)      # non-symmetric dressed Fock
    out = {"nocc": o, "nvir": v, "fsp_ns": fsp_ns}
    cc = CCSD.GCC(er)
    for fname, fsp in (("sym", synth.fsp(o, v)), ("ns", fsp_ns)):
        for tag, alpha, eq in MODES:
            a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha, equation=eq)
            out["T1_%s_%s" % (fname, tag)], out["T2_%s_%s" % (fname, tag)] = a, b
            a, b = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha, equation=eq)
            out["L1_%s_%s" % (fname, tag)], out["L2_%s_%s" % (fname, tag)] = a, b
        out["E_%s" % fname] = cc.energy(t1, t2, fsp)
    out["gamma"] = cc.gamma(t1, t2, l1, l2)
    # the reference's second implementation (CCSD.py:675-699 identity), bare Fock
    out["rawT1"], out["rawT2"] = RAW.T1T2eq(t1, t2, er)
    out["rawL1"], out["rawL2"] = RAW.La1La2eq(t1, t2, l1, l2, er)
    return out


def ccs_inputs(o, v):
    """Random amplitudes, non-symmetric dressed Fock and Vexp (cf. the reference's own check, CCS.py:2632-2639)."""
    n = o + v
    rng = np.random.default_rng(77 * o + v)
    d = {k: 0.1 * rng.standard_normal((o, v)) for k in ("ts", "ls", "rs", "rl")}
    d["fsp"] = synth.fock(o, v) + 0.05 * rng.standard_normal((n, n))
    d["vm"] = 0.05 * rng.standard_normal((n, n))
    d["vm2"] = 0.05 * rng.standard_normal((n, n))
    d.update(r0=0.3, l0=0.2, Em=0.7)
    return d


def ccs_calls(cc, mod, d):
    """Every public CCS entry point, fresh intermediates per call."""
    ts, ls, rs, rl, fsp, vm, vm2 = (d[k] for k in ("ts", "ls", "rs", "rl", "fsp", "vm", "vm2"))
    r0, l0, Em = d["r0"], d["l0"], d["Em"]
    out = {}

    def put(name, val):
        if isinstance(val, (tuple, list)):
            for i, x in enumerate(val):
                out["%s_%d" % (name, i)] = np.asarray(x, dtype=float)
        else:
            out[name] = np.asarray(val, dtype=float)

    put("T1inter", cc.T1inter(ts, fsp))
    put("T1eq", cc.T1eq(ts, fsp))
    put("tsupdate", cc.tsupdate(ts, cc.T1inter(ts, fsp)))
    put("tsupdate_es", cc.tsupdate(ts, cc.T1inter(ts, fsp), [rs, rl], [r0, 0.1], [vm, vm2]))
    put("tsupdate_L1", np.array(cc.tsupdate_L1(ts, cc.T1inter(ts, fsp), 1e-3)))
    put("L1inter", cc.L1inter(ts, fsp))
    put("L1eq", cc.L1eq(ts, ls, fsp))
    put("lsupdate", cc.lsupdate(ts, ls, cc.L1inter(ts, fsp)))
    put("lsupdate_es", cc.lsupdate(ts, ls, cc.L1inter(ts, fsp), [rs], [rl], [r0], [l0], [vm]))
    put("lsupdate_L1", cc.lsupdate_L1(ls, cc.L1inter(ts, fsp), 1e-3))
    put("R1inter", cc.R1inter(ts, fsp, vm))
    put("R1inter_novm", cc.R1inter(ts, fsp, None))
    put("Extract_Em_r", cc.Extract_Em_r(rs, r0, cc.R1inter(ts, fsp, vm)))
    put("rsupdate", cc.rsupdate(rs, r0, cc.R1inter(ts, fsp, vm), Em))
    put("rsupdate_noforce", cc.rsupdate(rs, r0, cc.R1inter(ts, fsp, vm), Em, force_alpha=False))
    put("get_ov", cc.get_ov(ls, l0, rs, r0, (1, 2)))
    put("R1eq", cc.R1eq(rs, r0, cc.R1inter(ts, fsp, vm)))
    put("R0inter", cc.R0inter(ts, fsp, vm))
    put("r0update", cc.r0update(rs, r0, Em, cc.R0inter(ts, fsp, vm)))
    put("R0eq", cc.R0eq(rs, r0, cc.R0inter(ts, fsp, vm)))
    put("r0_fromE", cc.r0_fromE(Em, ts, rs, vm, fsp))
    put("es_L1inter", cc.es_L1inter(ts, fsp, vm))
    put("L0inter", cc.L0inter(ts, fsp, vm))
    put("Extract_Em_l", cc.Extract_Em_l(ls, l0, cc.es_L1inter(ts, fsp, vm)))
    put("es_lsupdate", cc.es_lsupdate(ls, l0, Em, cc.es_L1inter(ts, fsp, vm)))
    put("es_L1eq", cc.es_L1eq(ls, l0, cc.es_L1inter(ts, fsp, vm)))
    put("l0update", cc.l0update(ls, l0, Em, cc.L0inter(ts, fsp, vm)))
    put("L0eq", cc.L0eq(ls, l0, cc.L0inter(ts, fsp, vm)))
    put("l0_fromE", cc.l0_fromE(Em, ts, ls, vm, fsp))
    put("energy_ccs", cc.energy_ccs(ts, fsp))
    put("energy_ccs_es", cc.energy_ccs(ts, fsp, [rs], [r0], [vm]))
    put("gamma", mod.gamma_CCS(ts, ls))
    put("gamma_unsym", mod.gamma_unsym_CCS(ts, ls))
    put("gamma_es", mod.gamma_es_CCS(ts, ls, rs, r0, l0))
    put("gamma_tr", mod.gamma_tr_CCS(ts, ls, rs, r0, l0))
    put("gamma_es_gs", mod.gamma_es_CCS(ts, ls, None, None, None))
    return out


def ccs_case(o, v):
    CCS = ref_loader.load("CCS")
    er = synth.SynthEris(o, v)
    d = ccs_inputs(o, v)
    out = {"nocc": o, "nvir": v}
    out.update({"in_" + k: np.asarray(val, dtype=float) for k, val in d.items()})
    out.update(ccs_calls(CCS.Gccs(er), CCS, d))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for o, v in ((4, 6), (6, 9)):
        np.savez_compressed(os.path.join(OUT, "ccs_o%dv%d.npz" % (o, v)), **ccs_case(o, v))
        print("wrote ccs_o%dv%d.npz" % (o, v))
    for o, v in ((4, 6), (5, 8)):
        np.savez_compressed(os.path.join(OUT, "ccsd_o%dv%d.npz" % (o, v)), **ccsd_case(o, v))
        print("wrote ccsd_o%dv%d.npz" % (o, v))


if __name__ == "__main__":
    main()

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


def main():
    os.makedirs(OUT, exist_ok=True)
    for o, v in ((4, 6), (5, 8)):
        np.savez_compressed(os.path.join(OUT, "ccsd_o%dv%d.npz" % (o, v)), **ccsd_case(o, v))
        print("wrote ccsd_o%dv%d.npz" % (o, v))


if __name__ == "__main__":
    main()

"""Driver hooks: build() compiles every CUDA extension for sm_100a and imports the
package; smoke() runs one tiny invocation of the hot path on cuda:0 against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build():
    import ecw_cc_b200
    ecw_cc_b200.build(verbose=True)           # nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...
    assert b"sm_100a" in ecw_cc_b200.lib.ecw_version()
    # the checker: the oracle is numpy (nothing to compile); the reference is pure Python, so there
    # is no oracle/_ref binary.  When the reference tree is present, verify the oracle still pins to it.
    from oracle import ref_loader
    if ref_loader.available():
        subprocess.check_call([sys.executable, "-c",
                               "from oracle import ref_loader; ref_loader.load('CCSD','CCS','utilities')"],
                              cwd=ROOT)


def smoke():
    import numpy as np
    import torch
    import ecw_cc_b200
    from oracle import synth
    from oracle.ccsd_np import OracleGCC
    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    torch.cuda.set_device(0)
    o, v = 4, 6
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    fsp = synth.fsp(o, v)
    cc, orc = ecw_cc_b200.GCC(er), OracleGCC(er)
    worst = 0.0
    for alpha in (None, 1e-3):
        a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha)
        c, d = orc.tupdate(t1, t2, fsp=fsp, alpha=alpha)
        worst = max(worst, np.abs(a - c).max(), np.abs(b - d).max())
        a, b = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha)
        c, d = orc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha)
        worst = max(worst, np.abs(a - c).max(), np.abs(b - d).max())
    worst = max(worst, np.abs(cc.gamma(t1, t2, l1, l2) - orc.gamma(t1, t2, l1, l2)).max())
    worst = max(worst, abs(cc.energy(t1, t2, fsp) - orc.energy(t1, t2, fsp)))
    assert worst < 1e-10, worst
    print("smoke ok: CCSD T/Lambda/gamma/energy on cuda:0 vs oracle, max abs diff %.2e" % worst)


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()

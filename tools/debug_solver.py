import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ecw_cc_b200 as ecw
from oracle import synth
from oracle.ccsd_np import OracleGCC
o, v = 5, 9
er = synth.SynthEris(o, v); fsp = synth.fsp(o, v)
e = np.diagonal(er.fock); d1 = e[:o, None] - e[None, o:]; d2 = d1[:, None, :, None] + d1[None, :, None, :]
cc, orc = ecw.GCC(er), OracleGCC(er)
ts = np.zeros((o, v)); ls = np.zeros((o, v)); td = er.oovv / d2; ld = td.copy()
for alpha in (None, 5e-4):
    for eq in (True, False):
        a, b = cc.tupdate(ts, td, fsp=fsp, alpha=alpha, equation=eq); c, d = orc.tupdate(ts, td, fsp=fsp, alpha=alpha, equation=eq)
        print('T alpha', alpha, 'eq', eq, np.abs(a - c).max(), np.abs(b - d).max())
        a2, b2 = cc.lupdate(c, d, ls, ld, fsp=fsp, alpha=alpha, equation=eq); c2, d2_ = orc.lupdate(c, d, ls, ld, fsp=fsp, alpha=alpha, equation=eq)
        print('L alpha', alpha, 'eq', eq, np.abs(a2 - c2).max(), np.abs(b2 - d2_).max())
        if not eq and alpha is not None:
            bad = np.argwhere(np.abs(b2 - d2_) > 1e-12)
            print('n bad L2', len(bad), bad[:5])
            bad = np.argwhere(np.abs(b - d) > 1e-12); print('n bad T2', len(bad), bad[:5])
            for ix in bad[:3]:
                ix = tuple(ix); print(ix, b[ix], d[ix], td[ix])

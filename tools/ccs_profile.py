"""Host-side profile (cProfile) of one ECW-CCS ground-state iteration body and one excited-state iteration body through
ecw_cc_b200.Gccs at a synthetic (nocc, nvir).  Usage: python tools/ccs_profile.py [nocc nvir]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ecw_cc_b200 as ecw

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (24, 240)
de = ecw.DeviceEris.synthetic(o, v, gemm="int8", keep_fp64_vvvv=True)
cc = ecw.Gccs(de)
n = o + v
rng = np.random.default_rng(3)
ts, ls, rs, rl = (0.05 * rng.standard_normal((o, v)) for _ in range(4))
fsp = de.fock + 0.02 * rng.standard_normal((n, n))
vm = 0.02 * rng.standard_normal((n, n))


def gs():
    t = cc.tsupdate(ts, cc.T1inter(ts, fsp))
    l = cc.lsupdate(t, ls, cc.L1inter(t, fsp))
    cc.gamma(t, l)
    cc.energy_ccs(t, fsp)


def es():
    ri = cc.R1inter(ts, fsp, vm)
    em, _, _ = cc.Extract_Em_r(rs, 0.3, ri)
    cc.rsupdate(rs, 0.3, ri, em)
    cc.r0update(rs, 0.3, em, cc.R0inter(ts, fsp, vm))
    li = cc.es_L1inter(ts, fsp, vm)
    el, _, _ = cc.Extract_Em_l(rl, 0.2, li)
    cc.es_lsupdate(rl, 0.2, el, li)
    cc.l0update(rl, 0.2, el, cc.L0inter(ts, fsp, vm))


for name, fn in (("gs", gs), ("es", es)):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print("%s: %.1f ms per iteration" % (name, 1e3 * (time.perf_counter() - t0) / 5))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

"""Where an iteration of ecw_cc_b200.Solver_CCSD.SCF spends its time at (40,400) (bring-up tool)."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ecw_cc_b200 as ecw
from ecw_cc_b200.exp_pot import Exp
o, v = 40, 400
n = o + v
de = ecw.DeviceEris.synthetic(o, v)
cc = ecw.GCC(de)
t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
l1, l2 = de.synth_tensor("l1", (o, v)), de.synth_tensor("l2", (o, o, v, v))
fsp = de.synth_tensor("fsp", (n, n))
def T(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("gamma", T(lambda: cc.gamma(t1, t2, l1, l2)))
print("gamma+cpu", T(lambda: cc.gamma(t1, t2, l1, l2).cpu().numpy()))
print("energy float", T(lambda: float(cc.energy(t1, t2, fsp))))
print("tupdate", T(lambda: cc.tupdate(t1, t2, fsp=fsp)))
print("lupdate", T(lambda: cc.lupdate(t1, t2, l1, l2, fsp=fsp)))
print("antisym check", T(lambda: cc.antisym_defect(t2)))
rng = np.random.default_rng(7)
pert = 0.02 * rng.standard_normal((n, n))
target = np.diag(np.concatenate([np.ones(o), np.zeros(v)])) + 0.5 * (pert + pert.T)
s = ecw.Solver_CCSD(cc, Exp(0.05, [[["mat", target]]]), conv_thres=0.0, maxiter=1)      # MP2 start
print("SCF 2 iterations", T(lambda: s.SCF(0.05, return_device=True), reps=2))
cn = torch.empty(t1.numel() + t2.numel(), dtype=torch.float64, device="cuda"); co = torch.zeros_like(cn)
print("conv", T(lambda: s._conv_distance(cn, co, t1, l1, t2, l2)))
# where does an SCF iteration spend its time?  (synchronising timers around every call of the loop)
acc = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    setattr(obj, name, g)
for nm in ("gamma", "energy", "tupdate", "lupdate", "antisym_defect"):
    wrap(cc, nm)
wrap(s, "_conv_distance")
torch.cuda.synchronize(); t0 = time.perf_counter()
out = s.SCF(0.05, return_device=True)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
print("SCF total %.1f ms for %d iterations; inside calls: %s; other %.1f" % (tot, len(out[1]), {k: round(x, 1) for k, x in acc.items()}, tot - sum(x for k, x in acc.items() if k != "antisym_defect")))
print("final (defect, max|x|): t2", cc.antisym_stats(out[5][2]), "l2", cc.antisym_stats(out[5][3]), "Ep", out[1])

"""Smallest program that launches the hot kernels at the benchmark shape, for ncu:
one `GCC.tupdate` at (nocc, nvir) = (40, 400) on synthetic integrals (INT8 engine by default).
    python tools/ncu_target.py [int8|dmma] [nocc nvir]
Order of the ozaki_gemm_kernel launches in a packed-path tupdate (csrc/ccsd_plan_slab.cpp, one GPU): 0 cc_Fvv tau~
(split-K), 1 T1 ovvv (batched), 2 Woooo tau.oovv, 3 hh ladder, 4 K1 pp ladder (79800 x 780 x 79800), 5 R9, 6 [ijae,be->ijab],
7 t1.ovvv (batched), 8 R1 Wovvo 16000^3, 9 R2 ring 16000^3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecw_cc_b200 as ecw

gemm = sys.argv[1] if len(sys.argv) > 1 else None
o, v = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40, 400)
de = ecw.DeviceEris.synthetic(o, v, gemm=gemm)
cc = ecw.GCC(de, assume_antisym=True)
n = o + v
t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
fsp = de.synth_tensor("fsp", (n, n))
a, b = cc.tupdate(t1, t2, fsp=fsp)
torch.cuda.synchronize()
print("tupdate done: |t1new| %.6e |t2new| %.6e" % (float(a.norm()), float(b.norm())))

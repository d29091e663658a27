"""Parity of the two GEMM engines AT THE BENCHMARK SHAPE: one T / Lambda / rdm1 / energy evaluation at
(nocc, nvir) = (40, 400) with the FP64 DMMA engine and with the INT8 tensor-core engine (digit planes, batched and
split-K products, constant ovvv/vvvv plane sets) on the same synthetic inputs; prints max |difference| per output.
The oracle cannot run this size (600 GB); the DMMA engine is itself pinned to the oracle at every size it can.
    python tools/engine_check.py [nocc nvir] [out.json]
"""
import gc
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecw_cc_b200 as ecw

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 400)
n = o + v
res = {}
for eng in ("dmma", "int8"):
    de = ecw.DeviceEris.synthetic(o, v, gemm=eng)
    cc = ecw.GCC(de)
    t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
    l1, l2 = de.synth_tensor("l1", (o, v)), de.synth_tensor("l2", (o, o, v, v))
    fsp = de.synth_tensor("fsp", (n, n))
    out = {}
    for alpha, tag in ((None, "upd"), (1e-3, "l1upd")):
        a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha)
        c, d = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha)
        for k, x in (("t1new_" + tag, a), ("t2new_" + tag, b), ("l1new_" + tag, c), ("l2new_" + tag, d)):
            out[k] = x.cpu()
    out["gamma"] = cc.gamma(t1, t2, l1, l2).cpu()
    out["energy"] = torch.tensor([float(cc.energy(t1, t2, fsp))], dtype=torch.float64)
    res[eng] = out
    print(eng, "digits", de.int8_digits, "peak mem %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
    del cc, de, t1, t2, l1, l2, fsp, a, b, c, d
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
report = {"nocc": o, "nvir": v}
for k in res["dmma"]:
    x, y = res["dmma"][k], res["int8"][k]
    report[k] = {"max_abs_diff": float((x - y).abs().max()), "max_abs": float(x.abs().max())}
    print("%-14s max|dmma - int8| = %.3e   (max|value| %.3e)" % (k, report[k]["max_abs_diff"], report[k]["max_abs"]), flush=True)
worst = max(r["max_abs_diff"] for r in report.values() if isinstance(r, dict))
print("worst %.3e  (bar: 1e-10)" % worst)
if len(sys.argv) > 3 or (len(sys.argv) == 2):
    json.dump(report, open(sys.argv[-1], "w"), indent=1)
assert worst < 1e-10

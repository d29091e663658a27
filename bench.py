#!/usr/bin/env python
"""bench.py — CCSD T+Lambda residual evaluations per second on B200.

One *step* = the work of one `Solver_CCSD.SCF` iteration body (Solver_GS.py:683-705):
`gamma` + `energy` + `tupdate` + `lupdate` on the synthetic spin-orbital workload
(nocc, nvir) = (40, 400), FP64 (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference algorithm on host cores

Prints ONE JSON line.  `value` = steps/s with everything resident in HBM; `e2e` = the same through
the reference-facing API (`GCC.gamma/energy/tupdate/lupdate`) with HOST (pinned numpy) buffers,
host<->device copies inside the timed region.  `roofline` is for the dominant launch (the packed
particle-particle ladder GEMM), timed live with CUDA events on the launching stream.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ccsd_t_lambda_residual_evals_per_sec"
UNIT = "evals/s"
NOMINAL_FP64_TFLOPS = 37.0     # B200 datasheet FP64 / FP64-tensor (context only)


def f_alg(o, v):
    """Algorithmic flops per evaluation."""
    po, pv = o * (o - 1) // 2, v * (v - 1) // 2
    return 4.0 * po * pv * pv + 18.0 * o ** 3 * v ** 3 + 14.0 * o ** 4 * v ** 2


def f_ref(o, v):
    """Dense flops the reference factorisation executes (K1-K5, R1-R7, o^4v^2 terms)."""
    return 8.0 * o ** 2 * v ** 4 + 2.0 * o * v ** 4 + 14.0 * o ** 3 * v ** 3 + 14.0 * o ** 4 * v ** 2


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode()
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU (reference algorithm)
def cpu_eval_time(o, v, reps=2, warm=1):
    """Seconds per evaluation of the reference algorithm (oracle port, reference's own einsum routing)."""
    from oracle import synth
    from oracle.ccsd_np import OracleGCC
    er = synth.SynthEris(o, v)
    t1, t2, l1, l2 = synth.amplitudes(o, v)
    fsp = synth.fsp(o, v)
    cc = OracleGCC(er, faithful=True)

    def one():
        cc.gamma(t1, t2, l1, l2)
        cc.energy(t1, t2, fsp)
        cc.tupdate(t1, t2, fsp=fsp)
        cc.lupdate(t1, t2, l1, l2, fsp=fsp)

    for _ in range(warm):
        one()
    best = 1e99
    for _ in range(reps):
        t0 = time.perf_counter()
        one()
        best = min(best, time.perf_counter() - t0)
    return best


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(o, v, sample=(12, 64), reps=2, warm=1):
    so, sv = sample
    sec = cpu_eval_time(so, sv, reps=reps, warm=warm)
    scaled = (1.0 / sec) * f_ref(so, sv) / f_ref(o, v)
    return {"value": scaled, "unit": UNIT, "cores": blas_threads(), "kind": "port",
            "sample": "reference algorithm (oracle/ccsd_np.py, reference's einsum routing) at (nocc,nvir)=(%d,%d): "
                      "%.3f s/eval measured = %.4g evals/s at that size; value is that rate scaled by the "
                      "reference's dense flop count to (%d,%d) (x%.3g) - the reference cannot run (%d,%d) itself "
                      "(~600 GB of v^4 intermediates)" % (so, sv, sec, 1.0 / sec, o, v,
                                                          f_ref(so, sv) / f_ref(o, v), o, v),
            "sample_evals_per_sec": 1.0 / sec, "sample_shape": [so, sv]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o, v = args.nocc, args.nvir
    sample = (12, 64)
    t = []
    for _ in range(min(args.warmup, 1)):        # one warm-up pass is enough for numpy; keeps the run bounded
        cpu_eval_time(sample[0], sample[1], reps=1, warm=0)
    for _ in range(args.steps):
        t.append(cpu_eval_time(sample[0], sample[1], reps=1, warm=0))
    sec = sum(t) / len(t)
    scaled = (1.0 / sec) * f_ref(*sample) / f_ref(o, v)
    cb = {"value": scaled, "unit": UNIT, "cores": blas_threads(), "kind": "port",
          "sample": "reference algorithm (oracle port, reference's einsum routing) at (%d,%d): %.3f s/eval, "
                    "scaled by the reference's dense flop count to (%d,%d)" % (sample[0], sample[1], sec, o, v),
          "sample_evals_per_sec": 1.0 / sec}
    line = {"impl": "reference", "metric": METRIC, "value": scaled, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / scaled, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic spin-orbital CCSD T+Lambda residual, nocc=%d nvir=%d FP64" % (o, v),
                       "nocc": o, "nvir": v, "sample_shape": list(sample)},
            "cpu_baseline": cb,
            "e2e": {"value": scaled, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- ours
def measure_fp64_peak(torch, n=8192, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e99
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, b, out=c)
        e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9


def ncu_traffic():
    """dram bytes read+written by the dominant launch, from the committed ncu --set full capture."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_int8_oz_tupdate_ncu.json")))
        rec = [r for r in d["launches"] if "pp ladder" in r.get("launch", "")][0]
        rd, wr = rec["dram__bytes_read.sum"].split(), rec["dram__bytes_write.sum"].split()
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        return {"bytes_per_launch": float(rd[0]) * unit[rd[1]] + float(wr[0]) * unit[wr[1]],
                "algorithmic_bytes_per_launch": 3.86e10 + 0.5e9,
                "source": "profiles/r1_int8_oz_tupdate_ncu.json (ncu --set full, same launch, 1 GPU)"}
    except Exception:
        return None


def time_parts(cc, timed, t1, t2, l1, l2, fsp, reps=2, alpha=1e-3):
    """ms per call of the four functions of one evaluation, and of the two updates with the L1 term (SURVEY §8d:
    "report also tupdate and lupdate separately, and with/without alpha").  `timed(fn, steps, warmup) -> (ms, out)`."""
    calls = [("gamma", lambda: cc.gamma(t1, t2, l1, l2)),
             ("energy", lambda: cc.energy(t1, t2, fsp)),
             ("tupdate", lambda: cc.tupdate(t1, t2, fsp=fsp, alpha=None)),
             ("lupdate", lambda: cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=None)),
             ("tupdate_alpha", lambda: cc.tupdate(t1, t2, fsp=fsp, alpha=alpha)),
             ("lupdate_alpha", lambda: cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha))]
    out = {"alpha": alpha, "reps": reps}
    for name, fn in calls:
        ms, res = timed(fn, reps, 1)
        del res
        out[name + "_ms"] = ms / reps
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ecw_cc_b200 as ecw
    from ecw_cc_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    o, v = args.nocc, args.nvir
    n = o + v

    de = ecw.DeviceEris.synthetic(o, v, rank=rank, world=world, gemm=args.gemm, int8_digits=args.int8_digits)
    cc = ecw.GCC(de)
    ns = de.int8_digits
    t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
    l1, l2 = de.synth_tensor("l1", (o, v)), de.synth_tensor("l2", (o, o, v, v))
    fsp = de.synth_tensor("fsp", (n, n))
    alpha = args.alpha

    def step_dev():
        g = cc.gamma(t1, t2, l1, l2)
        e = cc.energy(t1, t2, fsp)
        a, b = cc.tupdate(t1, t2, fsp=fsp, alpha=alpha)
        c, d = cc.lupdate(t1, t2, l1, l2, fsp=fsp, alpha=alpha)
        return g, e, a, b, c, d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            out = fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, out

    # ---- device-resident leg (value) with clocks sampled during the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.ecw_profile_enable(de._h, 0)
    ms_dev, _ = timed(step_dev, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the functions of one evaluation on their own, and the updates with the L1 term (extra key "parts")
    try:
        parts = time_parts(cc, timed, t1, t2, l1, l2, fsp)
    except Exception as exc:                                   # never lose the bench line over the breakdown
        parts = {"error": repr(exc)[:200]}

    # ---- dominant launch, timed live with CUDA events on the launching stream
    lib.ecw_profile_enable(de._h, 1)
    ladder_ms = []
    for _ in range(min(args.steps, 3)):
        cc.tupdate(t1, t2, fsp=fsp, alpha=alpha)
        buf = ctypes.create_string_buffer(1 << 22)
        lib.ecw_profile_dump(de._h, buf, 1 << 22)
        ops = json.loads(buf.value.decode())
        ladder_ms += [x["ms"] for x in ops if x["kind"] in ("gemm", "oz_gemm") and "K1 pp ladder" in x["note"]]
    lib.ecw_profile_enable(de._h, 0)
    ladder = [x for x in ops if x["kind"] in ("gemm", "oz_gemm") and "K1 pp ladder" in x["note"]][0]
    ladder_flops = 2.0 * ladder["M"] * ladder["N"] * ladder["K"]
    ladder_ms = sum(ladder_ms) / len(ladder_ms)
    t_ms = sum(x["ms"] for x in ops)
    gemm_ms = sum(x["ms"] for x in ops if x["kind"] in ("gemm", "oz_gemm", "oz_split"))
    int8_ms = sum(x["ms"] for x in ops if x["kind"] in ("oz_gemm", "oz_split"))

    # ---- end-to-end leg: host (pinned numpy) buffers through the reference-facing API
    cc.h2d_bytes = cc.d2h_bytes = 0
    h = {k: cc._to_host(x) for k, x in (("t1", t1), ("t2", t2), ("l1", l1), ("l2", l2), ("fsp", fsp))}

    def step_host():
        g = cc.gamma(h["t1"], h["t2"], h["l1"], h["l2"])
        e = cc.energy(h["t1"], h["t2"], h["fsp"])
        a, b = cc.tupdate(h["t1"], h["t2"], fsp=h["fsp"], alpha=alpha)
        c, d = cc.lupdate(h["t1"], h["t2"], h["l1"], h["l2"], fsp=h["fsp"], alpha=alpha)
        return g, e, a, b, c, d

    e2e_steps = max(1, min(args.steps, 3))
    step_host()
    cc.h2d_bytes = cc.d2h_bytes = 0
    ms_e2e, _ = timed(step_host, e2e_steps, 1)
    h2d = cc.h2d_bytes // (e2e_steps + 1)
    d2h = cc.d2h_bytes // (e2e_steps + 1)

    # ---- the solver loop itself, device resident (ecw_cc_b200.Solver_CCSD mirrors Solver_GS.Solver_CCSD.SCF): per
    # iteration gamma -> Vexp('mat') on the host (n x n) -> energy -> tupdate -> lupdate -> convergence vector
    import numpy as np
    from ecw_cc_b200.exp_pot import Exp
    rng = np.random.default_rng(7)
    pert = 0.02 * rng.standard_normal((n, n))
    target = np.diag(np.concatenate([np.ones(o), np.zeros(v)])) + 0.5 * (pert + pert.T)
    # One iteration per SCF call (maxiter=0) from the synthetic amplitudes: the random integrals of this size are far
    # from a perturbative system (E_MP2 = -361), so the reference's quasi-Newton iteration blows up after an update and
    # would leave the antisymmetric (packed) path the value/e2e legs are measured on.
    solver = ecw.Solver_CCSD(cc, Exp(0.05, [[["mat", target]]]), conv_thres=0.0, maxiter=0,
                             tsini=t1, lsini=l1, tdini=t2, ldini=l2)
    solver_calls = max(1, min(args.steps, 3))

    def solver_run():
        return solver.SCF(0.05, alpha=alpha, return_device=True)

    ms_solver, sout = timed(solver_run, solver_calls, 1)
    solver_iters = solver_calls

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    anti = alpha is None
    launches = sum(cc.plan_launches(f, alpha, False, anti) for f in ("tupdate", "lupdate"))
    launches += cc.plan_launches("gamma") + cc.plan_launches("energy") + 3   # + antisymmetry checks
    exec_flops = (cc.plan_flops("tupdate", alpha, False, anti) + cc.plan_flops("lupdate", alpha, False, anti)
                  + cc.plan_flops("gamma") + cc.plan_flops("energy"))
    peak = measure_fp64_peak(torch)
    evals_per_s = args.steps * world / (ms_dev / 1e3) if world == 1 else args.steps / (ms_dev / 1e3)
    e2e_per_s = e2e_steps / (ms_e2e / 1e3)
    fp64_equiv = ladder_flops / ladder_ms / 1e9
    if ladder["kind"] == "oz_gemm":
        # the dominant launch runs on the INT8 tensor pipe: ns(ns+1)/2 int8 products per FP64 product
        nprod = ns * (ns + 1) // 2
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = ncu_traffic() if (o, v) == (40, 400) else None    # the committed capture is of the default workload
        # the ladder is timed inside the step (per-op events between back-to-back launches): the SUSTAINED figure is the
        # denominator the measurement rules name for that; the burst-based fraction is reported next to it
        bf16_burst = float(mp.get("bf16_tflops", 0.0))
        bf16 = float(mp.get("bf16_tflops_sustained", 0.0)) or bf16_burst
        int8_peak = 2.0 * bf16 if bf16 > 0 else 4500.0
        roofline = {"bound": "tensor",
                    "kernel": "ecw::ozaki_gemm_kernel<%d> (tcgen05.mma kind::i8, TMEM accumulators), launch = packed "
                              "pp-ladder %dx%dx%d (CCSD.py:305)" % (ns, ladder["M"], ladder["N"], ladder["K"]),
                    "achieved": fp64_equiv * nprod, "peak": int8_peak, "unit": "TOP/s",
                    "unit_note": "dense int8 tensor ops (2 per multiply-add): the TFLOP/s of a kernel whose operands "
                                 "are int8 digits; the FP64-equivalent rate is in fp64_equivalent_tflops",
                    "frac": fp64_equiv * nprod / int8_peak,
                    "traffic": (traffic or {}).get("bytes_per_launch"), "traffic_detail": traffic,
                    "frac_vs_burst_peak": (fp64_equiv * nprod / (2.0 * bf16_burst)) if bf16_burst > 0 else None,
                    "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops_sustained (%.1f; kernel timed inside a long "
                                    "step; burst figure %.1f): the INT8 dense rate of the tcgen05 pipe is twice the "
                                    "bf16 rate (nominal 4500 vs 2250)" % (bf16, bf16_burst)) if bf16 > 0
                    else "nominal INT8 dense 4500 TOP/s (B200_PROFILING.md fallback)",
                    "int8_products_per_fp64_product": nprod,
                    "fp64_equivalent_tflops": fp64_equiv, "cublas_dgemm_tflops_measured": peak,
                    "fp64_equivalent_over_fp64_tensor_peak": fp64_equiv / peak,
                    "launch_ms": ladder_ms, "launch_flops": ladder_flops,
                    "gemm_share_of_tupdate": gemm_ms / t_ms, "int8_share_of_tupdate": int8_ms / t_ms}
    else:
        roofline = {"bound": "tensor",
                    "kernel": "ecw::dgemm_tma_kernel (FP64 DMMA), launch = packed pp-ladder %dx%dx%d (CCSD.py:305)"
                              % (ladder["M"], ladder["N"], ladder["K"]),
                    "achieved": fp64_equiv, "peak": peak, "unit": "TFLOP/s",
                    "frac": fp64_equiv / peak, "traffic": None,
                    "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run, burst; "
                                   "MEASURED_PEAKS.json has no FP64 entry; nominal FP64 tensor peak %.0f TFLOP/s"
                                   % NOMINAL_FP64_TFLOPS,
                    "launch_ms": ladder_ms, "launch_flops": ladder_flops,
                    "gemm_share_of_tupdate": gemm_ms / t_ms}
    line = {
        "metric": METRIC, "value": evals_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic spin-orbital CCSD T+Lambda residual (gamma+energy+tupdate+lupdate), "
                               "nocc=%d nvir=%d FP64" % (o, v),
                   "nocc": o, "nvir": v, "alpha": alpha, "parallelism": "1 GPU" if world == 1 else "vshard%d" % world,
                   "gemm_engine": ("int8 tcgen05 (%d digits) for the large GEMMs, FP64 DMMA for the rest" % ns) if ns
                   else "FP64 DMMA",
                   "l2_policy": "inputs larger than L2 (the packed vvvv, %.1f GB, is streamed every step)"
                                % ((ns if ns else 8) * (v * (v - 1) // 2) ** 2 / 1e9),
                   "fp64_tensor_peak_measured_tflops": peak,
                   "tflops_alg_over_fp64_tensor_peak": f_alg(o, v) * evals_per_s / 1e12 / peak,
                   "f_alg_flops_per_eval": f_alg(o, v), "executed_gemm_flops_per_eval": exec_flops,
                   "tflops_alg": f_alg(o, v) * evals_per_s / 1e12, "tflops_executed": exec_flops * evals_per_s / 1e12},
        "clocks": clocks,
        "e2e": {"value": e2e_per_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "solver_loop": {"value": solver_iters / (ms_solver / 1e3), "unit": "iterations/s",
                        "h2d_bytes_per_iteration": 0, "d2h_bytes_per_iteration": 32,
                        "what": "ecw_cc_b200.Solver_CCSD.SCF (mirror of Solver_GS.Solver_CCSD.SCF) with "
                                "ecw_cc_b200.exp_pot.Exp ('mat' target): the same iteration with the amplitudes, the "
                                "rdm1, the experimental potential and the dressed Fock resident on the GPU; only "
                                "Delta, vmax, the energy and the convergence distance (4 doubles) cross PCIe"},
        "parts": parts,
        "gpu_launches": int(launches * args.steps),
        "roofline": roofline,
    }
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(o, v)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nocc", type=int, default=40)
    ap.add_argument("--nvir", type=int, default=400)
    ap.add_argument("--alpha", type=float, default=None, help="L1 coefficient (default: none, as Main.CCSD_GS)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gemm", default=None, choices=["int8", "dmma"], help="GEMM engine (default: int8)")
    ap.add_argument("--int8-digits", type=int, default=None, help="base-256 digits of the INT8 engine (default: from the integral magnitudes, 6 for the benchmark)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

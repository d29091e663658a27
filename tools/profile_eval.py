"""Per-op device timing of one CCSD T+Lambda residual evaluation (CUDA events around every
launch of the plan).  Usage: python tools/profile_eval.py NOCC NVIR [out.json] [int8|dmma] [digits]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecw_cc_b200 as ecw
from ecw_cc_b200 import lib

o, v = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time()
gemm = sys.argv[4] if len(sys.argv) > 4 else None
digits = int(sys.argv[5]) if len(sys.argv) > 5 else None
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:                        # torchrun: one rank per GPU, rank 0 reports (its ops include the waits on the others)
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
de = ecw.DeviceEris.synthetic(o, v, gemm=gemm, int8_digits=digits, rank=rank, world=world)
print("ranks: %d" % world)
print("gemm engine: %s" % ("int8, %d digits" % de.int8_digits if de.int8_digits else "dmma"), flush=True)
torch.cuda.synchronize()
print("eris ready %.1fs, mem %.1f GB" % (time.time() - t0, torch.cuda.memory_allocated() / 1e9), flush=True)
cc = ecw.GCC(de)
n = o + v
t1, t2 = de.synth_tensor("t1", (o, v)), de.synth_tensor("t2", (o, o, v, v))
l1, l2 = de.synth_tensor("l1", (o, v)), de.synth_tensor("l2", (o, o, v, v))
fsp = de.synth_tensor("fsp", (n, n))
res = {}
for name, fn in (("tupdate", lambda: cc.tupdate(t1, t2, fsp=fsp)), ("lupdate", lambda: cc.lupdate(t1, t2, l1, l2, fsp=fsp)),
                 ("gamma", lambda: cc.gamma(t1, t2, l1, l2))):
    fn(); torch.cuda.synchronize()            # warm
    lib.ecw_profile_enable(de._h, 1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); e.synchronize()
    buf = ctypes.create_string_buffer(1 << 22)
    r = lib.ecw_profile_dump(de._h, buf, 1 << 22)
    lib.ecw_profile_enable(de._h, 0)
    ops = json.loads(buf.value.decode())
    tot = sum(x["ms"] for x in ops)
    print("%s: %.1f ms wall(events) %.1f ms sum-of-ops, flops %.3e -> %.1f TFLOP/s" % (
        name, s.elapsed_time(e), tot, cc.plan_flops(name), cc.plan_flops(name) / max(s.elapsed_time(e), 1e-9) / 1e9), flush=True)
    bykind = {}
    for x in ops:
        bykind[x["kind"]] = bykind.get(x["kind"], 0.0) + x["ms"]
    print("   by kind:", {k: round(val, 1) for k, val in sorted(bykind.items(), key=lambda kv: -kv[1])})
    for x in sorted(ops, key=lambda x: -x["ms"])[:22]:
        fl = 2.0 * x["M"] * x["N"] * x["K"] * x["batch"] / max(x["splitk"], 1) if x["kind"] in ("gemm", "oz_gemm") else 0
        print("   %8.2f ms %-8s M%-7d N%-7d K%-8d b%-5d %6.1f TF  %s" % (x["ms"], x["kind"], x["M"], x["N"], x["K"], x["batch"], fl / max(x["ms"], 1e-9) / 1e9, x["note"]))
    res[name] = {"ms": s.elapsed_time(e), "ops": ops}
print("peak mem %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
if len(sys.argv) > 3 and sys.argv[3] != "-" and rank == 0:
    json.dump(res, open(sys.argv[3], "w"))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""Kernel experiments for csrc/ozaki.cu (bring-up tool): where does the INT8 GEMM lose tensor-pipe time?
ECW_OZ_DEBUG=1 makes the producer skip the loads (MMA issue rate on stale shared memory),
ECW_OZ_GRID=n limits the persistent grid to n CTAs."""
import os, sys, json, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecw_cc_b200
from ecw_cc_b200 import lib

st = torch.cuda.current_stream().cuda_stream
ns = 6
M = N = K = 16000
A = torch.randn((M, K), dtype=torch.float64, device="cuda") * 0.02
B = torch.randn((N, K), dtype=torch.float64, device="cuda") * 0.01
C = torch.empty((M, N), dtype=torch.float64, device="cuda")


def cut(X):
    R, Kx = X.shape
    planes = torch.empty(lib.ecw_ozaki_plane_bytes(R, Kx, ns), dtype=torch.int8, device="cuda")
    stat = torch.empty(lib.ecw_ozaki_stat_elems(R), dtype=torch.float64, device="cuda")
    assert lib.ecw_ozaki_split(X.data_ptr(), R, Kx, X.stride(0), 1, ns, planes.data_ptr(), stat.data_ptr(), st) == 0
    return planes, stat


pa, sa = cut(A)
pb, sb = cut(B)


def run(reps=4):
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        assert lib.ecw_ozaki_gemm(pa.data_ptr(), sa.data_ptr(), pb.data_ptr(), sb.data_ptr(), M, N, K, C.data_ptr(), 1, M,
                                  1.0, 0.0, ns, st) == 0
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return ts


def clock():
    out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"]).decode()
    return out.strip()


# ECW_OZ_DEBUG bits: 1 no loads, 2 A planes only, 4 B planes only, 16 only 15 of the 21 products
for name, env in [("normal", {}), ("noload", {"ECW_OZ_DEBUG": "1"}), ("grid74", {"ECW_OZ_GRID": "74"}),
                  ("grid74_noload", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "1"}),
                  ("grid74_Aonly", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "2"}),
                  ("grid74_Bonly", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "4"}),
                  ("grid74_15prod", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "16"}),
                  ("grid74_15prod_noload", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "17"}),
                  ("grid74_15prod_Aonly", {"ECW_OZ_GRID": "74", "ECW_OZ_DEBUG": "18"})]:
    for k in ("ECW_OZ_DEBUG", "ECW_OZ_GRID"):
        os.environ.pop(k, None)
    os.environ.update(env)
    run(1)
    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append(clock())
            time.sleep(0.05)
    th = threading.Thread(target=sampler)
    th.start()
    ts = run(6)
    stop.set()
    th.join()
    fl = 2.0 * M * N * K
    print(name, "ms", ["%.1f" % t for t in ts], "TF %.1f" % (fl / min(ts) / 1e9), "clk/power", samples[len(samples) // 2:][:3], flush=True)

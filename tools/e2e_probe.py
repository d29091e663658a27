"""Where does the host side of a numpy-API call go?  Pinned allocation vs the D2H / H2D copies (2 GB arrays)."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
n = 40 * 40 * 400 * 400
d = torch.randn(n, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
def T(f, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return ["%.1f ms" % (1e3 * x) for x in ts], r
print("pinned alloc 2 GB (kept alive):", T(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True))[0])
keep = []
print("pinned alloc 2 GB, previous kept:", T(lambda: keep.append(torch.empty(n, dtype=torch.float64, pin_memory=True)))[0])
h = keep[0]
print("D2H 2 GB into pinned:", T(lambda: h.copy_(d, non_blocking=True))[0])
print("H2D 2 GB from pinned:", T(lambda: d.copy_(h, non_blocking=True))[0])
p = torch.empty(n, dtype=torch.float64)
print("D2H 2 GB into pageable:", T(lambda: p.copy_(d))[0])
a = h.numpy()
print("numpy view + setflags:", T(lambda: a.setflags(write=False))[0])

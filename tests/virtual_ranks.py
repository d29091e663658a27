"""Virtual ranks: several sharded contexts of the product in ONE process on ONE GPU, one Python thread per rank, the
collectives of `DeviceEris.comm` emulated by device copies between the ranks' workspaces.  Lets `pytest -m gpu` on a
single B200 run the code every rank of a multi-GPU job runs (ecw_ctx_set_shard, contract_lead_dist, contract_split,
ladder_dist, the owner-computes ring GEMMs) without NCCL.  All threads launch on the same (default) stream, so device
work is ordered by launch time; the barriers order the launches."""
import threading


class ThreadComm(object):
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.send = [None] * world
        self.vals = [None] * world
        self.gathers = 0

    def all_gather(self, eris, recv, send):
        r, n = eris.rank, send.numel()
        self.send[r] = send
        self.barrier.wait()                      # every rank has launched what produces its contribution
        for q in range(self.world):
            recv[q * n:(q + 1) * n].copy_(self.send[q])
        if r == 0:
            self.gathers += 1
        self.barrier.wait()                      # nobody overwrites a contribution before everyone has copied it

    def all_to_all(self, eris, recv, send):
        r, n = eris.rank, send.numel() // self.world
        self.send[r] = send
        self.barrier.wait()
        for q in range(self.world):
            recv[q * n:(q + 1) * n].copy_(self.send[q][r * n:(r + 1) * n])
        if r == 0:
            self.gathers += 1
        self.barrier.wait()

    def max_scalar(self, eris, x):
        self.vals[eris.rank] = float(x.cpu()[0])
        self.barrier.wait()
        v = max(self.vals)
        self.barrier.wait()
        return v


def run_ranks(world, body):
    """body(rank, comm) -> result, run on `world` threads; returns the list of results (re-raises the first error)."""
    import torch
    comm = ThreadComm(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            torch.cuda.set_device(0)
            out[r] = body(r, comm)
        except BaseException as e:               # noqa: BLE001 - reported in the main thread
            err[r] = e
            comm.barrier.abort()

    ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return out, comm

"""Host-side mirror of the reference `CCSD.GCC` class (CCSD.py:185-623).

Same constructor and method signatures as the reference so that the unchanged
solver loop (`Solver_GS.Solver_CCSD.SCF`, Solver_GS.py:621-742) drives the CUDA
kernels:  `gamma`, `energy`, `tupdate`, `lupdate`.  numpy arrays in -> fresh
numpy arrays out (inputs are never mutated); torch CUDA tensors in -> torch
CUDA tensors out (device-resident mode used by bench.py's `value` leg).
"""
import numpy as np

from ._lib import lib, EcwError, ECW_HAS_ALPHA, ECW_EQUATION, ECW_ANTISYM
from .eris import DeviceEris


# Amplitudes whose antisymmetry defect max|x[ijab]+x[jiab]|, |x[ijab]+x[ijba]| is below this run
# the fully packed path (the defect only enters the result multiplied by O(1) integrals/amplitudes,
# far below the 1e-10 parity bar); anything larger — e.g. after an L1-regularised update, which
# breaks antisymmetry in the reference (utilities.py:59-67) — runs the general path.
ANTISYM_TOL = 1e-13


def _flags(alpha, equation, antisym=False):
    return ((ECW_HAS_ALPHA if alpha is not None else 0) | (ECW_EQUATION if equation else 0)
            | (ECW_ANTISYM if antisym else 0))


class GCC(object):
    def __init__(self, eris, fock=None, device=None, assume_antisym=None, rank=0, world=1, group=None):
        """:param eris: a `DeviceEris`, or any object with the reference's
        `Eris.geris` attribute surface (uploaded once).
        :param assume_antisym: None (default) = measure the antisymmetry of the doubles amplitudes
        on the device at every call and pick the packed or the general path; True/False = force."""
        self.assume_antisym = assume_antisym
        if not isinstance(eris, DeviceEris):
            eris = DeviceEris.from_geris(eris, device=device, rank=rank, world=world, group=group)
        self.eris = eris
        self.nocc = eris.nocc
        if fock is None:                       # CCSD.py:196-198
            self.fock = eris.fock
        self.nvir = self.fock.shape[0] - self.nocc
        self._pin = {}
        self.h2d_bytes = 0      # bytes staged host->device / device->host by this object
        self.d2h_bytes = 0

    # -- host <-> device staging ---------------------------------------------
    def _torch(self):
        import torch
        return torch

    def _to_dev(self, name, x, shape):
        """numpy -> device.  Arrays that already live in pinned host memory (e.g. the arrays this
        class returns) are DMA'd directly; others are staged through a persistent pinned buffer."""
        torch = self._torch()
        if isinstance(x, torch.Tensor):
            if x.dtype != torch.float64 or tuple(x.shape) != tuple(shape):
                raise ValueError("%s: expected float64 tensor of shape %s" % (name, (shape,)))
            return x.contiguous(), True
        a = np.asarray(x, dtype=np.float64)
        if a.shape != tuple(shape):
            raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), a.shape))
        self.h2d_bytes += a.nbytes
        if a.flags.c_contiguous and a.flags.writeable:
            t = torch.from_numpy(a)
            if t.is_pinned():
                return t.to(self.eris.device, non_blocking=True), False
        key = (name, tuple(shape))
        pin = self._pin.get(key)
        if pin is None:
            pin = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            self._pin[key] = pin
        else:
            torch.cuda.current_stream(self.eris.device).synchronize()   # previous DMA out of this buffer
        pin.numpy()[...] = a
        return pin.to(self.eris.device, non_blocking=True), False

    def _to_host(self, *ts):
        """device -> fresh numpy arrays backed by pinned memory (torch's caching host allocator)."""
        torch = self._torch()
        outs = []
        for t in ts:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            self.d2h_bytes += t.numel() * 8
            outs.append(h)
        torch.cuda.current_stream(self.eris.device).synchronize()
        outs = [h.numpy() for h in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)

    def antisym_defect(self, d_x):
        """max |x[ijab]+x[jiab]|, |x[ijab]+x[ijba]| of a device doubles amplitude."""
        torch = self._torch()
        e = self.eris
        out = torch.zeros(1, dtype=torch.float64, device=e.device)
        if lib.ecw_antisym_defect(d_x.data_ptr(), self.nocc, self.nvir, out.data_ptr(), e.stream()) != 0:
            raise EcwError("ecw_antisym_defect failed")
        return float(out.cpu()[0])

    def _is_antisym(self, *amps):
        if self.assume_antisym is not None:
            return bool(self.assume_antisym)
        return all(self.antisym_defect(a) <= ANTISYM_TOL for a in amps)

    def _fsp(self, fsp):
        n = self.nocc + self.nvir
        if fsp is None:
            return self.eris.fock_dev, True
        return self._to_dev("fsp", fsp, (n, n))

    # -- rdm1 (CCSD.py:204-208 -> :136-182) -------------------------------------
    def gamma(self, t1, t2, l1, l2):
        torch = self._torch()
        o, v = self.nocc, self.nvir
        e = self.eris
        d_t1, dev = self._to_dev("t1", t1, (o, v))
        d_t2, _ = self._to_dev("t2", t2, (o, o, v, v))
        d_l1, _ = self._to_dev("l1", l1, (o, v))
        d_l2, _ = self._to_dev("l2", l2, (o, o, v, v))
        out = torch.empty((o + v, o + v), dtype=torch.float64, device=e.device)
        e.ensure_workspace("gamma", 0)
        e.run(lib.ecw_ccsd_gamma(e._h, d_t1.data_ptr(), d_t2.data_ptr(), d_l1.data_ptr(), d_l2.data_ptr(),
                                 out.data_ptr(), e.stream()), "ecw_ccsd_gamma")
        return out if dev else self._to_host(out)

    # -- energy (CCSD.py:224-242) -------------------------------------------------
    def energy(self, t1, t2, fsp):
        torch = self._torch()
        o, v = self.nocc, self.nvir
        e = self.eris
        d_t1, dev = self._to_dev("t1", t1, (o, v))
        d_t2, _ = self._to_dev("t2", t2, (o, o, v, v))
        d_f, _ = self._fsp(fsp)
        out = torch.empty(1, dtype=torch.float64, device=e.device)
        e.ensure_workspace("energy", 0)
        e.run(lib.ecw_ccsd_energy(e._h, d_t1.data_ptr(), d_t2.data_ptr(), d_f.data_ptr(), out.data_ptr(),
                                  e.stream()), "ecw_ccsd_energy")
        if dev:
            return out[0]
        self.d2h_bytes += 8
        return float(out.cpu()[0])

    # -- T1/T2 (CCSD.py:248-338) ---------------------------------------------------
    def tupdate(self, t1, t2, fsp=None, alpha=None, equation=False):
        torch = self._torch()
        o, v = self.nocc, self.nvir
        e = self.eris
        d_t1, dev = self._to_dev("t1", t1, (o, v))
        d_t2, _ = self._to_dev("t2", t2, (o, o, v, v))
        d_f, _ = self._fsp(fsp)
        o1 = torch.empty((o, v), dtype=torch.float64, device=e.device)
        o2 = torch.empty((o, o, v, v), dtype=torch.float64, device=e.device)
        fl = _flags(alpha, equation, self._is_antisym(d_t2))
        e.ensure_workspace("tupdate", fl)
        e.run(lib.ecw_ccsd_tupdate(e._h, d_t1.data_ptr(), d_t2.data_ptr(), d_f.data_ptr(), e.fock_dev.data_ptr(),
                                   fl, float(alpha or 0.0), o1.data_ptr(), o2.data_ptr(), e.stream()),
              "ecw_ccsd_tupdate")
        if dev:
            return o1, o2
        return self._to_host(o1, o2)

    # -- L1/L2 (CCSD.py:419-535, Linter :543-623) --------------------------------------
    def lupdate(self, t1, t2, l1, l2, fsp=None, alpha=None, equation=False):
        torch = self._torch()
        o, v = self.nocc, self.nvir
        e = self.eris
        d_t1, dev = self._to_dev("t1", t1, (o, v))
        d_t2, _ = self._to_dev("t2", t2, (o, o, v, v))
        d_l1, _ = self._to_dev("l1", l1, (o, v))
        d_l2, _ = self._to_dev("l2", l2, (o, o, v, v))
        d_f, _ = self._fsp(fsp)
        o1 = torch.empty((o, v), dtype=torch.float64, device=e.device)
        o2 = torch.empty((o, o, v, v), dtype=torch.float64, device=e.device)
        fl = _flags(alpha, equation, self._is_antisym(d_t2, d_l2))
        e.ensure_workspace("lupdate", fl)
        e.run(lib.ecw_ccsd_lupdate(e._h, d_t1.data_ptr(), d_t2.data_ptr(), d_l1.data_ptr(), d_l2.data_ptr(),
                                   d_f.data_ptr(), e.fock_dev.data_ptr(), fl, float(alpha or 0.0),
                                   o1.data_ptr(), o2.data_ptr(), e.stream()), "ecw_ccsd_lupdate")
        if dev:
            return o1, o2
        return self._to_host(o1, o2)

    # -- introspection ---------------------------------------------------------------
    def plan_json(self, func, alpha=None, equation=False, antisym=True):
        import ctypes
        fl = _flags(alpha, equation, antisym)
        n = 1 << 22
        while True:
            buf = ctypes.create_string_buffer(n)
            r = lib.ecw_plan_dump(self.eris._h, func.encode(), fl, buf, n)
            if r >= 0:
                return buf.value.decode()
            if r == -1:
                self.eris.check(-1, "ecw_plan_dump")
            n = int(-r) + 16

    def plan_flops(self, func, alpha=None, equation=False, antisym=True):
        return lib.ecw_plan_flops(self.eris._h, func.encode(), _flags(alpha, equation, antisym))

    def plan_launches(self, func, alpha=None, equation=False, antisym=True):
        return lib.ecw_plan_launches(self.eris._h, func.encode(), _flags(alpha, equation, antisym))


def gamma_CCSD(t1, t2, l1, l2, mycc=None):
    """Module-level rdm1 of the reference (CCSD.py:136-162); needs a GCC for the device context."""
    if mycc is None:
        raise EcwError("gamma_CCSD needs the owning GCC object (device context)")
    return mycc.gamma(t1, t2, l1, l2)

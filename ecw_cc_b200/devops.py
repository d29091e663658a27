"""Thin host wrapper over the primitive device ops of the C ABI (`ecw_op_*`,
include/ecw_b200.h): einsum-style contraction, axpby/permutes, element-wise
product, diagonal shift, denominators and dot products on torch CUDA tensors.
Torch only allocates the memory and provides the stream; every arithmetic
operation is one of this repo's kernels.
"""
import ctypes

import numpy as np

from ._lib import lib, EcwError


class _TensorDesc(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("nd", ctypes.c_int32),
                ("dim", ctypes.c_int64 * 6), ("str", ctypes.c_int64 * 6)]


def _desc(t):
    if t is None:
        return None
    d = _TensorDesc()
    d.ptr = t.data_ptr()
    d.nd = t.dim()
    for i in range(t.dim()):
        d.dim[i] = t.shape[i]
        d.str[i] = t.stride(i)
    return d


def _scalar(x):
    """Python float from a number or a one-element array (the solvers pass r0 / l0 / E as shape-(1,) arrays)."""
    return float(np.asarray(x, dtype=np.float64).reshape(-1)[0])


def _ref(d):
    return ctypes.byref(d) if d is not None else None


class DevOps(object):
    def __init__(self, eris):
        import torch
        self.torch = torch
        self.e = eris
        self.dev = eris.device
        self.launches = 0

    # -- memory -----------------------------------------------------------------
    def to_dev(self, x):
        torch = self.torch
        if isinstance(x, torch.Tensor):
            return x
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.dev)

    def empty(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device=self.dev)

    POOL_MIN_BYTES = 1 << 20    # larger downloads go through pooled pinned blocks (a pageable D2H runs at ~2.5 GB/s)
    POOL_DEPTH = 3

    def to_host(self, t):
        """device -> fresh numpy array.  Large results (the o^2v^2 intermediates Wbija / Wakic) are DMA'd into pinned
        blocks that return to a per-shape pool when the caller drops the array, as in GCC._to_host."""
        if t.numel() * 8 < self.POOL_MIN_BYTES:
            return t.cpu().numpy()
        import weakref
        torch = self.torch
        pool = self.__dict__.setdefault("_pool", {})
        key = tuple(t.shape)
        free = pool.setdefault(key, [])
        if free:
            h = free.pop()
            del free[self.POOL_DEPTH:]
        else:
            h = torch.empty(key, dtype=torch.float64, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        a = h.numpy()
        weakref.finalize(a, free.append, h)
        return a

    def scalar(self, t):
        """Device scalar -> Python float (one synchronising D2H read)."""
        return float(t.cpu().reshape(-1)[0])

    # -- dispatch with workspace growth ---------------------------------------------
    def _call(self, fn, *args):
        e = self.e
        if e._ws is None:
            e._ws = self.torch.empty(1 << 20, dtype=self.torch.uint8, device=self.dev)
            e.check(lib.ecw_set_workspace(e._h, e._ws.data_ptr(), e._ws.numel()), "ecw_set_workspace")
        rc = fn(e._h, *args, e.stream())
        if rc == -2:
            need = int(lib.ecw_op_workspace_needed(e._h))
            self.torch.cuda.current_stream(self.dev).synchronize()
            e._ws = None
            e._ws = self.torch.empty(max(need, 256) + 4096, dtype=self.torch.uint8, device=self.dev)
            e.check(lib.ecw_set_workspace(e._h, e._ws.data_ptr(), e._ws.numel()), "ecw_set_workspace")
            rc = fn(e._h, *args, e.stream())
        e.check(rc, getattr(fn, "__name__", "ecw_op"))
        self.launches += 1

    # -- ops ---------------------------------------------------------------------------
    def contract(self, spec, A, B, alpha=1.0, out=None, beta=0.0):
        """out[sc] = alpha * einsum(spec, A, B) + beta * out.  spec like 'jb,jabi->ai'."""
        lhs, sc = spec.split("->")
        sa, sb = lhs.split(",")
        if out is None:
            dims = {}
            for s, t in ((sa, A), (sb, B)):
                for ch, n in zip(s, t.shape):
                    dims[ch] = n
            out = self.empty(*[dims[ch] for ch in sc])
            beta = 0.0
        da, db, dc = _desc(A), _desc(B), _desc(out)
        args = (_scalar(alpha), _ref(da), sa.encode(), _ref(db), sb.encode(), _scalar(beta), _ref(dc), sc.encode())
        self._call(lib.ecw_op_contract, *args)
        e = self.e
        if e.int8_digits and e.int8_tol > 0.0 and 2.0 * A.numel() * B.numel() >= e.int8_min_flops >= 0.0:
            # a contraction this large may have taken the INT8 route (2MNK <= 2|A||B|): accuracy guard (eris.py)
            b = e.last_bound = e.int8_bound()
            if b == b and b > e.int8_tol:
                e.guard_trips += 1
                if _scalar(beta) != 0.0:
                    raise EcwError("ecw_op_contract: INT8 error bound %.2e > %.1e on an accumulating contraction; "
                                   "use more digits (int8_digits) or gemm='dmma'" % (b, e.int8_tol))
                e.check(lib.ecw_ctx_set_engine_override(e._h, 1), "ecw_ctx_set_engine_override")
                try:
                    self._call(lib.ecw_op_contract, *args)
                finally:
                    lib.ecw_ctx_set_engine_override(e._h, 0)
        return out

    def axpby(self, alpha, A, sa, beta, C, sc):
        da, dc = _desc(A), _desc(C)
        self._call(lib.ecw_op_axpby, _scalar(alpha), _ref(da), sa.encode(), _scalar(beta), _ref(dc), sc.encode())
        return C

    def copy(self, A, spec=None, alpha=1.0):
        """Fresh contiguous tensor = alpha * A (optionally permuted: spec 'ai->ia')."""
        lab = "pqrstu"[: A.dim()]
        sa, sc = (spec.split("->") if spec else (lab, lab))
        dims = dict(zip(sa, A.shape))
        out = self.empty(*[dims[ch] for ch in sc])
        return self.axpby(alpha, A, sa, 0.0, out, sc)

    def add(self, C, A, alpha=1.0, spec=None):
        """C += alpha * A (optionally permuted)."""
        lab = "pqrstu"[: A.dim()]
        sa, sc = (spec.split("->") if spec else (lab, lab))
        return self.axpby(alpha, A, sa, 1.0, C, sc)

    def mul(self, alpha, A, B, beta, C):
        da, db, dc = _desc(A), _desc(B), _desc(C)
        self._call(lib.ecw_op_mul, _scalar(alpha), _ref(da), _ref(db), _scalar(beta), _ref(dc))
        return C

    def fill(self, C, value):
        return self.mul(value, None, None, 0.0, C)

    def scale_add(self, C, A, alpha):
        """C += alpha * A (same shape, any strides)."""
        return self.mul(alpha, A, None, 1.0, C)

    def unpack(self, A2, flags, C4, alpha=1.0, beta=0.0):
        da, dc = _desc(A2), _desc(C4)
        self._call(lib.ecw_op_unpack, _scalar(alpha), _ref(da), int(flags), _scalar(beta), _ref(dc))
        return C4

    def diag_shift(self, C, alpha, offset):
        dc, df = _desc(C), _desc(self.e.fock_dev)
        self._call(lib.ecw_op_diag_shift, _ref(dc), _scalar(alpha), _ref(df), int(offset))
        return C

    def denom(self, resid, amp, flags=0, alpha=0.0, shift=0.0, out=None):
        if out is None:
            out = self.empty(*resid.shape)
        dr, da, df, do = _desc(resid), _desc(amp), _desc(self.e.fock_dev), _desc(out)
        self._call(lib.ecw_op_denom, _ref(dr), _ref(da), _ref(df), int(self.e.nocc), int(flags), _scalar(alpha),
                   _scalar(shift), _ref(do))
        return out

    def dot(self, A, B, alpha=1.0):
        """alpha * <A, B> as a Python float."""
        out = self.torch.zeros(1, dtype=self.torch.float64, device=self.dev)
        da, db = _desc(A), _desc(B)
        self._call(lib.ecw_op_dot, _scalar(alpha), _ref(da), _ref(db), 0.0, ctypes.c_void_p(out.data_ptr()))
        return self.scalar(out)

    def trace(self, M):
        """sum_i M[i,i] through the dot kernel (diagonal view x ones)."""
        n = M.shape[0]
        diag = self.torch.as_strided(M, (n,), (M.stride(0) + M.stride(1),))
        ones = self.fill(self.empty(n), 1.0)
        return self.dot(diag, ones)

"""numpy interpreter for execution plans — TEST INFRASTRUCTURE ONLY.

Replays the op list that `ecw_plan_dump` emits (the same list `exec.cu`
launches on the GPU) with numpy on host arrays, so the *lowering* — index
strings, layouts, batch/split-K choices, workspace offsets and lifetimes — can
be checked against the oracle on a CPU-only box.  It is not a fallback: the
product has no path that reaches it.
"""
import json

import numpy as np
from numpy.lib.stride_tricks import as_strided


def npair(n):
    return n * (n - 1) // 2


def pair_decode(n):
    hi, lo = np.tril_indices(n, -1)
    return lo, hi


# ---- INT8-pipe GEMM (csrc/ozaki.cu): numpy statement of the digit cut and the device plane order
ZERO_ROW_SCALE = 2.0 ** -600


def oz_scale(X):
    """power-of-two row scales s_r > max|X[r,:]| (2^-600 for an all-zero row: its products underflow to exact zeros)."""
    mx = np.abs(X).max(axis=1) if X.shape[1] else np.zeros(X.shape[0])
    _, e = np.frexp(mx)
    return np.where(mx > 0, np.ldexp(1.0, e), ZERO_ROW_SCALE)


def oz_cprime(ns):
    return sum(256.0 ** (-q) for q in range(1, ns))


def oz_digits(X, ns):
    """X[r,k] = s_r (2 sum_p d_p 256^-(p+1) + c) + delta: int8 digits [ns][R][K] and the scales.
    Y = min(rint(y 256^ns), 256^ns - 1) with y = (x/s + 1)/2; digit p is byte ns-1-p of Y, minus 128."""
    s = oz_scale(X)
    y = X * (0.5 / s)[:, None] + 0.5                 # one rounding, as the device fma
    z = np.rint(y * 2.0 ** (8 * ns))                 # exact scaling; ties to even like cvt.rn
    top = 2.0 ** (8 * ns)
    Y = np.where(z >= top, 0.0, z).astype(np.uint64)
    Y = np.where(z >= top, np.uint64(2 ** (8 * ns) - 1), Y)
    out = [(((Y >> np.uint64(8 * (ns - 1 - p))) & np.uint64(255)).astype(np.int64) - 128).astype(np.int8)
           for p in range(ns)]
    return np.stack(out), s


def oz_value(D, s, ns):
    """the FP64 numbers the digits stand for."""
    Dv = sum(D[p].astype(np.float64) * 256.0 ** (-(p + 1)) for p in range(ns))
    return s[:, None] * (2.0 * Dv + oz_cprime(ns))


def oz_to_planes(D, K1=1):
    """digits [ns][R][K1*K2] -> device order [kb][p][rg][j][ri][16]: rows padded to 128, k2 to 32 per k1 (flat int8)."""
    ns, R, K = D.shape
    K2 = K // K1
    Rp, K2p = (R + 127) // 128 * 128, (K2 + 31) // 32 * 32
    P = np.zeros((ns, Rp, K1, K2p), dtype=np.int8)
    P[:, :R, :, :K2] = D.reshape(ns, R, K1, K2)
    P = P.reshape(ns, Rp, K1 * K2p)
    return np.ascontiguousarray(P.reshape(ns, Rp // 8, 8, K1 * K2p // 32, 2, 16).transpose(3, 0, 1, 4, 2, 5)).reshape(-1)


def oz_from_planes(buf, ns, Rp, nkb):
    """flat int8 -> digits [ns][Rp][nkb*32] (padded rows and k kept)."""
    return buf[: ns * Rp * nkb * 32].reshape(nkb, ns, Rp // 8, 2, 8, 16).transpose(1, 2, 4, 0, 3, 5).reshape(ns, Rp, nkb * 32)


def oz_stats(X, s, K1=1):
    """device statistics array [row scales (padded to 128) | row sums / scale | the same per k1 (K1 > 1)]."""
    R = X.shape[0]
    Rp = (R + 127) // 128 * 128
    st = np.zeros((2 + K1 if K1 > 1 else 2) * Rp)
    st[:Rp] = ZERO_ROW_SCALE
    st[:R] = s
    st[Rp: Rp + R] = X.sum(axis=1) / s
    if K1 > 1:
        part = X.reshape(R, K1, -1).sum(axis=2) / s[:, None]
        for k1 in range(K1):
            st[(2 + k1) * Rp: (2 + k1) * Rp + R] = part[:, k1]
    return st


def oz_const_slots(X, ns, K1=1):
    """(planes, stats) slot arrays for a constant operand X[R,K] (what ecw_eris_*_planes leaves bound)."""
    D, s = oz_digits(X, ns)
    pl = oz_to_planes(D, K1)
    pl = np.concatenate([pl, np.zeros(4096 + (-pl.size) % 8, np.int8)]).view(np.float64).copy()
    return pl, oz_stats(X, s, K1)


def oz_product(DA, sa, ta, DB, sb, tb, K, ns):
    """what ozaki_gemm_kernel forms from two digit sets (exact integer products, FP64 Horner)."""
    DA, DB = DA.astype(np.float64), DB.astype(np.float64)
    H = np.zeros((DA.shape[1], DB.shape[1]))
    for w in range(ns - 1, -1, -1):
        part = np.zeros_like(H)
        for p in range(w + 1):
            part += DA[p] @ DB[w - p].T               # exact: integers far below 2^53
        H = H * 0.00390625 + part
    c = oz_cprime(ns)
    return sa[:, None] * sb[None, :] * (H * 6.103515625e-05 + (c * ta - c * c * K)[:, None] + (c * tb)[None, :])


class Interp(object):
    def __init__(self, plan_json, slots, alpha=0.0, allgather=None):
        self.allgather = allgather      # callable(send ndarray, recv ndarray) for multi-rank plans
        self.plan = json.loads(plan_json) if isinstance(plan_json, str) else plan_json
        self.slots = dict(slots)
        self.alpha = alpha
        ws = int(self.plan["workspace_elems"])
        # poison the workspace so reads of never-written scratch show up as NaN
        self.slots["ws"] = np.full(max(ws, 1), np.nan)
        if "scal" not in self.slots:
            self.slots["scal"] = np.zeros(16)

    def view(self, t):
        if t is None:
            return None
        base = self.slots[t["slot"]].reshape(-1)
        dim = tuple(t["dim"])
        strides = tuple(8 * s for s in t["str"])
        off = t["off"]
        # bounds check
        ext = off + sum((d - 1) * s for d, s in zip(t["dim"], t["str"]) if d > 0)
        assert 0 <= off and ext < base.size or np.prod(dim) == 0, (t, base.size)
        return as_strided(base[off:], shape=dim, strides=strides)

    def run(self):
        for op in self.plan["ops"]:
            getattr(self, "op_" + op["kind"])(op)
        return self

    # ---------------------------------------------------------------- ops
    def _mat(self, t_off_slot, off, rows, cols, rs, cs):
        base = self.slots[t_off_slot].reshape(-1)
        return as_strided(base[off:], shape=(rows, cols), strides=(8 * rs, 8 * cs))

    def op_gemm(self, op):
        M, N, K = op["M"], op["N"], op["K"]
        S = op["splitk"]
        kc = op["kchunk"] if S > 1 else K
        for zb in range(op["batch"]):
            r, s = divmod(zb, S)
            k0 = s * kc
            kn = min(kc, K - k0)
            assert kn > 0
            a, b, c = op["a"], op["b"], op["c"]
            if op["ta"] == 0:
                A = self._mat(a["slot"], a["off"] + r * op["sA"] + k0, M, kn, op["lda"], 1)
            else:
                A = self._mat(a["slot"], a["off"] + r * op["sA"] + k0 * op["lda"], M, kn, 1, op["lda"])
            if op["tb"] == 0:
                B = self._mat(b["slot"], b["off"] + r * op["sB"] + k0 * op["ldb"], kn, N, op["ldb"], 1)
            else:
                B = self._mat(b["slot"], b["off"] + r * op["sB"] + k0, kn, N, 1, op["ldb"])
            C = self._mat(c["slot"], c["off"] + zb * op["sC"], M, N, op["ldc"], 1)
            assert not np.isnan(A).any() and not np.isnan(B).any(), op["note"]
            res = op["alpha"] * (A @ B)
            if op["beta"] != 0.0:
                res = res + op["beta"] * C
            C[...] = res

    def op_reduce(self, op):
        M, N = op["M"], op["N"]
        a, c = op["a"], op["c"]
        P = self._mat(a["slot"], a["off"], op["i0"], M * N, M * N, 1).reshape(op["i0"], M, N)
        C = self._mat(c["slot"], c["off"], M, N, op["i1"], op["i2"])
        res = op["alpha"] * P.sum(axis=0)
        if op["beta"] != 0.0:
            res = res + op["beta"] * C
        C[...] = res

    def op_permute(self, op):
        A, C = self.view(op["a"]), self.view(op["c"])
        assert not np.isnan(A).any(), op["note"]
        res = op["alpha"] * A
        if op["beta"] != 0.0:
            res = res + op["beta"] * C
        C[...] = res

    def op_allgather(self, op):
        ws = self.slots["ws"]
        count, world = op["i0"], op["i1"]
        send = ws[op["a"]["off"]: op["a"]["off"] + count]
        recv = ws[op["c"]["off"]: op["c"]["off"] + world * count]
        assert self.allgather is not None, "multi-rank plan needs an allgather callable"
        self.allgather(send, recv)

    def op_alltoall(self, op):
        """emulated through the all-gather callable (test infrastructure: every rank sees every send buffer)"""
        ws = self.slots["ws"]
        count, world, rank = op["i0"], op["i1"], op["i2"]
        send = ws[op["a"]["off"]: op["a"]["off"] + world * count]
        allsend = np.empty(world * world * count)
        assert self.allgather is not None, "multi-rank plan needs an allgather callable"
        self.allgather(send, allsend)
        recv = ws[op["c"]["off"]: op["c"]["off"] + world * count]
        for q in range(world):
            recv[q * count:(q + 1) * count] = allsend[(q * world + rank) * count:(q * world + rank + 1) * count]

    def op_asym4(self, op):
        z, out = self.view(op["a"]), self.view(op["c"])
        assert not np.isnan(z).any(), op["note"]
        res = z - z.transpose(1, 0, 2, 3) - z.transpose(0, 1, 3, 2) + z.transpose(1, 0, 3, 2)
        if op["b"] is not None and op["alpha"] != 0.0:
            res = res + op["alpha"] * self.view(op["b"])
        if op["beta"] != 0.0:
            res = res + op["beta"] * out
        out[...] = res

    def _bytes(self, t):
        return self.slots[t["slot"]].reshape(-1).view(np.int8)[8 * t["off"]:]

    def op_oz_split(self, op):
        a = op["a"]
        R, K2, K1, ns = op["M"], op["K"], op["i1"], op["i0"]
        base = self.slots[a["slot"]].reshape(-1)
        X = np.array(as_strided(base[a["off"]:], shape=(R, K1, K2),
                                strides=(8 * op["lda"], 8 * op["ldc"], 8 * op["ldb"]))).reshape(R, K1 * K2)
        assert not np.isnan(X).any(), op["note"]
        D, s = oz_digits(X, ns)
        pl = oz_to_planes(D, K1)
        assert pl.size + 4096 <= 8 * op["c"]["dim"][0], op["note"]
        self._bytes(op["c"])[: pl.size] = pl
        st = self.view(op["d"])
        want = oz_stats(X, s, K1)
        assert st.shape[0] == want.size, op["note"]
        st[...] = want

    def op_oz_gemm(self, op):
        M, N, K, ns = op["M"], op["N"], op["K"], op["i0"]
        a_row0, a_rowb, b_row0, b_rowb, a_kb0, a_kbb, b_kb0, b_kbb, a_t0, a_tb, b_t0, b_tb, nkb = op["oz"]
        if nkb == 0:
            nkb = (K + 31) // 32
        Ap, Bp = (op["lda"] + 127) // 128 * 128, (op["ldb"] + 127) // 128 * 128
        sta, stb = self.view(op["d"]), self.view(op["e"])
        ba, bb = self._bytes(op["a"]), self._bytes(op["b"])
        nka, nkb_b = (8 * op["a"]["dim"][0] - 4096) // (ns * Ap * 32), (8 * op["b"]["dim"][0] - 4096) // (ns * Bp * 32)
        DA, DB = oz_from_planes(ba, ns, Ap, nka), oz_from_planes(bb, ns, Bp, nkb_b)
        c = op["c"]
        # the product must not write over the planes / statistics it reads (workspace lifetimes of the plan)
        c_lo = c["off"]
        c_hi = c["off"] + (op["batch"] - 1) * op["sC"] + (M - 1) * op["i1"] + (N - 1) * op["i2"]
        for k in "abde":
            t = op[k]
            if t and t["slot"] == c["slot"]:
                assert c_hi < t["off"] or c_lo > t["off"] + t["dim"][0] - 1, "oz_gemm output overlaps operand %s: %s" % (k, op["note"])
        for b in range(op["batch"]):
            ar, br = a_row0 + b * a_rowb, b_row0 + b * b_rowb
            ka, kb = (a_kb0 + b * a_kbb) * 32, (b_kb0 + b * b_kbb) * 32
            da = DA[:, ar:ar + M, ka:ka + nkb * 32]
            db = DB[:, br:br + N, kb:kb + nkb * 32]
            ta = sta[a_t0 + b * a_tb + ar: a_t0 + b * a_tb + ar + M]
            tb = stb[b_t0 + b * b_tb + br: b_t0 + b * b_tb + br + N]
            sa, sb = sta[ar:ar + M], stb[br:br + N]
            assert not (np.isnan(sa).any() or np.isnan(sb).any() or np.isnan(ta).any() or np.isnan(tb).any()), op["note"]
            C = self._mat(c["slot"], c["off"] + b * op["sC"], M, N, op["i1"], op["i2"])
            res = op["alpha"] * oz_product(da, sa, ta, db, sb, tb, K, ns)
            if op["beta"] != 0.0:
                res = res + op["beta"] * C
            C[...] = res

    def op_fill(self, op):
        self.view(op["c"])[...] = op["alpha"]

    def op_tau(self, op):
        t2, t1, out = self.view(op["a"]), self.view(op["b"]), self.view(op["c"])
        x = np.einsum('ia,jb->ijab', t1, t1)
        full = op["alpha"] * x - op["beta"] * x.transpose(0, 1, 3, 2)
        if op["i2"]:                                     # rows i0 .. i0+ni-1 of the leading occupied index
            full = full[op["i0"]: op["i0"] + op["i1"]]
        out[...] = t2 + full

    def op_pack(self, op):
        A, C = self.view(op["a"]), self.view(op["c"])
        fl = op["i0"]
        x = A
        if fl & 4:
            x = x - x.transpose(0, 1, 3, 2)
        if fl & 8:
            x = x - x.transpose(1, 0, 2, 3)
        if fl & 2:
            lo, hi = pair_decode(A.shape[2])
            x = x[:, :, lo, hi]
        else:
            x = x.reshape(A.shape[0], A.shape[1], -1)
        if fl & 1:
            lo, hi = pair_decode(A.shape[0])
            x = x[lo, hi]
        else:
            x = x.reshape(A.shape[0] * A.shape[1], -1)
        res = op["alpha"] * x
        if op["beta"] != 0.0:
            res = res + op["beta"] * C
        C[...] = res

    def op_unpack(self, op):
        A, C = self.view(op["a"]), self.view(op["c"])
        fl = op["i0"]
        d0, d1, d2, d3 = C.shape
        x = A
        if fl & 2:
            lo, hi = pair_decode(d2)
            y = np.zeros((x.shape[0], d2, d3))
            y[:, lo, hi] = x
            y[:, hi, lo] = -x
        else:
            y = x.reshape(x.shape[0], d2, d3)
        if fl & 1:
            lo, hi = pair_decode(d0)
            w = np.zeros((d0, d1, d2, d3))
            w[lo, hi] = y
            w[hi, lo] = -y
        else:
            w = y.reshape(d0, d1, d2, d3)
        res = op["alpha"] * w
        if op["beta"] != 0.0:
            res = res + op["beta"] * C
        C[...] = res

    def op_finish(self, op):
        r, amp, fock, out = self.view(op["a"]), self.view(op["b"]), self.view(op["d"]), self.view(op["c"])
        o, rank, has_alpha, equation = op["i0"], op["i1"], op["i2"], op["i3"]
        alpha = self.alpha
        e = np.diagonal(fock)
        d1 = e[:o, None] - e[None, o:]
        d = d1 if rank == 2 else d1[:, None, :, None] + d1[None, :, None, :]
        if has_alpha:
            if rank == 4:
                shr = np.where(r < -alpha, r + alpha, np.where(r > alpha, r - alpha, 0.0))
                w = np.where(amp > 0.0, r + alpha, shr)
            else:
                w = r
            res = w if equation else (w + amp * d) / d
        else:
            res = r if equation else r / d
        out[...] = res

    def op_dot(self, op):
        A, B = self.view(op["a"]), self.view(op["b"])
        k = op["i0"]
        s = self.slots["scal"]
        s[k] = (op["beta"] * s[k] if op["beta"] != 0.0 else 0.0) + op["alpha"] * float(np.sum(A * B))

    def op_scale_dev(self, op):
        C = self.view(op["c"])
        C *= op["d0"] + op["d1"] * self.slots["scal"][op["i0"]]

    def op_diag_add(self, op):
        C, fock = self.view(op["c"]), self.view(op["d"])
        m = C.shape[0]
        idx = np.arange(m)
        C[idx, idx] += op["alpha"] * np.diagonal(fock)[op["i0"]:op["i0"] + m]

    def op_rdm1(self, op):
        doo, dvoT, l1, dvv, out = (self.view(op[k]) for k in ("a", "b", "d", "e", "c"))
        o = doo.shape[0]
        out[:o, :o] = 0.5 * (doo + doo.T)
        out[:o, o:] = 0.5 * (l1 + dvoT)
        out[o:, :o] = out[:o, o:].T
        out[o:, o:] = 0.5 * (dvv + dvv.T)
        out[np.arange(o), np.arange(o)] += 1.0

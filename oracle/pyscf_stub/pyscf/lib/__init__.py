"""`pyscf.lib` subset used on the reference hot path (test infrastructure only).

`einsum`     : PySCF's lib.einsum is a BLAS-backed Einstein summation; numpy's
               optimize=True path (tensordot -> BLAS) is the same contraction up
               to FP64 summation order.
`direct_sum` : only the 'ia,jb->ijab' form occurs (CCSD.py:324,334,521,531;
               Solver_GS.py:557).
`diis`       : Pulay DIIS, used by the solver loops only when `diis` is set.
"""
import numpy as _np
from . import diis  # noqa: F401


def einsum(subscripts, *operands, **kwargs):
    return _np.einsum(subscripts, *operands, optimize=True)


def direct_sum(subscripts, *operands):
    if subscripts.replace(" ", "") != "ia,jb->ijab":
        raise NotImplementedError("pyscf stub: direct_sum(%r)" % subscripts)
    a, b = operands
    return a[:, None, :, None] + b[None, :, None, :]

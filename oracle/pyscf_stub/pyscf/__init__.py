"""Minimal stand-in for the `pyscf` package — TEST INFRASTRUCTURE ONLY.

PySCF is not installable in the build container (no network, no libcint).  The
reference hot path (`/root/reference/ECW_CC/{CCSD,CCS,utilities,CC_raw_equations}.py`)
touches only `pyscf.lib.einsum`, `pyscf.lib.direct_sum` and, for the raw
equations, `pyscf.ccn.util.p`; everything else is needed at import time only
(`utilities.py:13-16`).  This stub provides exactly those names so the
*unmodified* reference modules can be imported by `oracle/ref_loader.py` to
pin the numpy restatement and to generate the golden vectors under
`tests/golden/`.  It is never imported by the product package.
"""
from . import lib  # noqa: F401


class _Placeholder:
    """Import-time placeholder for sub-packages the hot path never calls."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        raise ImportError("pyscf stub: %s.%s is not available (hot-path-only stub)" % (self._name, item))


scf = _Placeholder("pyscf.scf")
tdscf = _Placeholder("pyscf.tdscf")
cc = _Placeholder("pyscf.cc")
ao2mo = _Placeholder("pyscf.ao2mo")
ci = _Placeholder("pyscf.ci")
from . import gto  # noqa: E402,F401

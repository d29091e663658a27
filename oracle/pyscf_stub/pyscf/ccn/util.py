"""`pyscf.ccn.util.p` subset used by CC_raw_equations.py:12 (CCS/CCSD sections
only): antisymmetrising permutation operator over the marked index pair."""


def p(spec, t):
    if spec == "ab..":
        return t - t.transpose(1, 0, 2, 3)
    if spec == "..ab":
        return t - t.transpose(0, 1, 3, 2)
    raise NotImplementedError("pyscf stub: p(%r)" % spec)

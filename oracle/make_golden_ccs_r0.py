"""Golden vectors of `Gccs.Extract_r0` (CCS.py:1036-1079): the UNMODIFIED reference on the synthetic CCS
inputs of oracle/make_golden.ccs_inputs, for several r1 vectors per size (the given one, scaled and sign-flipped
versions — both branches of the root selection and the ValueError).  Build container only:

    python -m oracle.make_golden_ccs_r0        -> tests/golden/ccs_extract_r0.npz
"""
import os
import warnings

import numpy as np

from . import ref_loader, synth
from .make_golden import OUT, ccs_inputs

SIZES = [(4, 6), (6, 9), (7, 12)]


def r1_variants(d):
    rs = d["rs"]
    return [rs, -rs, 3.0 * rs, 0.2 * rs + 0.1 * d["rl"], d["rl"], -0.5 * d["rl"] + 0.3 * d["ts"]]


def call(cc, r1, d, vm):
    """-> (value, status): status 0 = a number came back, 1 = ValueError('Both solution ...')."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return float(np.asarray(cc.Extract_r0(r1, d["ts"], d["fsp"], vm)).reshape(-1)[0]), 0
        except ValueError as e:
            assert "negative" in str(e)
            return np.nan, 1


def main():
    CCS = ref_loader.load("CCS")
    out = {}
    for o, v in SIZES:
        cc = CCS.Gccs(synth.SynthEris(o, v))
        d = ccs_inputs(o, v)
        vals, stat = [], []
        for vm in (d["vm"], d["vm2"]):
            for r1 in r1_variants(d):
                x, s = call(cc, r1, d, vm)
                vals.append(x)
                stat.append(s)
        x, s = call(cc, d["rs"], dict(d, fsp=None), d["vm"])          # fsp=None -> bare Fock (CCS.py:1044-1047)
        vals.append(x)
        stat.append(s)
        out["r0_o%dv%d" % (o, v)] = np.array(vals)
        out["status_o%dv%d" % (o, v)] = np.array(stat)
        print((o, v), np.array(vals), stat)
    np.savez_compressed(os.path.join(OUT, "ccs_extract_r0.npz"), **out)


if __name__ == "__main__":
    main()

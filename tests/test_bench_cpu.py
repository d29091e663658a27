"""Host logic of bench.py that can be checked without a GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_parts_breakdown_shape():
    import bench

    class FakeCC(object):
        def __init__(self):
            self.calls = []

        def gamma(self, *a):
            self.calls.append(("gamma", None))
            return 0

        def energy(self, *a):
            self.calls.append(("energy", None))
            return 0.0

        def tupdate(self, t1, t2, fsp=None, alpha=None):
            self.calls.append(("tupdate", alpha))
            return 1, 2

        def lupdate(self, t1, t2, l1, l2, fsp=None, alpha=None):
            self.calls.append(("lupdate", alpha))
            return 1, 2

    def timed(fn, steps, warmup):
        for _ in range(steps + warmup):
            out = fn()
        return 10.0 * steps, out

    cc = FakeCC()
    parts = bench.time_parts(cc, timed, 1, 2, 3, 4, 5, reps=2, alpha=1e-3)
    assert set(parts) == {"alpha", "reps", "gamma_ms", "energy_ms", "tupdate_ms", "lupdate_ms", "tupdate_alpha_ms",
                          "lupdate_alpha_ms"}
    assert all(parts[k] == 10.0 for k in parts if k.endswith("_ms"))
    assert cc.calls.count(("tupdate", None)) == 3 and cc.calls.count(("lupdate", 1e-3)) == 3


def test_ncu_traffic_reads_the_committed_capture():
    import bench
    t = bench.ncu_traffic()
    assert t is not None and t["bytes_per_launch"] > t["algorithmic_bytes_per_launch"] > 0

"""SASS evidence digest of the shipped library (profiles/r2_sass_excerpt.txt): per kernel, the counts of the Blackwell
instructions that matter (UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,
UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, DMMA = FP64 mma.sync) with one example line each.
Usage: python tools/sass_digest.py [libecw_b200.so] > profiles/r2_sass_excerpt.txt   (needs cuobjdump and c++filt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ecw_cc_b200", "libecw_b200.so")
KEYS = ("UTCIMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "DMMA")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts, example, total = None, {}, {}, collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern], example[kern] = collections.Counter(), {}
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1).split(".")[0]
    if op in KEYS:
        counts[kern][op] += 1
        total[op] += 1
        example[kern].setdefault(op, line.strip())
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence for the shipped library ecw_cc_b200/libecw_b200.so (cuobjdump -sass, sm_100a), round 2, final build.")
print("# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (linear bulk TMA),")
print("# UTMALDG = cp.async.bulk.tensor (tiled TMA), SYNCS = mbarrier operations, DMMA = FP64 mma.sync (m8n8k4).")
print("# Regenerate: python tools/sass_digest.py > profiles/r2_sass_excerpt.txt\n")
print("TOTAL over the library: " + ", ".join("%s x%d" % kv for kv in sorted(total.items())))
only_dmma = [k for k in counts if counts[k] and set(counts[k]) == {"DMMA"}]
print("kernels with only DMMA among these (the cp.async FP64 GEMM family ecw::dgemm_kernel<...>): %d\n" % len(only_dmma))
for k, nm in zip(counts, names):
    if not counts[k] or k in only_dmma:
        continue
    print("== " + nm)
    print("   " + ", ".join("%s x%d" % kv for kv in sorted(counts[k].items())))
    for op in ("UTMALDG", "UBLKCP", "LDTM", "UTCIMMA", "UTCBAR", "DMMA"):
        if op in example[k]:
            print("      " + example[k][op])
    print()

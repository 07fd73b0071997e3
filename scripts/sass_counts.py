"""SASS mnemonic counts of the kernels in the built library (evidence for profiles/: FFMA2 / LDS / UBLKCP (cp.async.bulk) /
UTCHMMA (tcgen05.mma) / LDTM (tcgen05.ld) / 256-bit LDG / STG ...).  Usage: python scripts/sass_counts.py [path to .so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "expressive_speech_synthesis_research_b200", "libwavernn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
fn, counts = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["instr"] += 1
        counts[fn][op.split(".")[0]] += 1
        if ".256" in op:
            counts[fn][op.split(".")[0] + ".256"] += 1
keys = ["instr", "FFMA2", "FFMA", "LDS", "STS", "SHFL", "LDG", "STG", "LDG.256", "STG.256", "UBLKCP", "SYNCS", "BAR", "UTCHMMA", "LDTM", "UTMALDG", "MUFU"]
print("%-52s " % "kernel" + " ".join("%8s" % k for k in keys))
for fn in sorted(counts):
    print("%-52s " % fn[:52] + " ".join("%8d" % counts[fn].get(k, 0) for k in keys))

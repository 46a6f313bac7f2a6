"""Opcode summary of the in-tree library per kernel (evidence that the kernels are Blackwell-native):
    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
Counts, per kernel of brats2019_b200/libbrats_b200.so, the SASS mnemonics that matter (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "brats2019_b200", "libbrats_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "HGMMA", "LDGSTS", "FFMA2", "REDUX"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                total[k] += 1
print("library: %s (sm_100a); %d kernels" % (os.path.relpath(so, REPO), len(counts)))
print("%-58s %6s " % ("kernel", "instr") + " ".join("%7s" % k for k in KEYS))
for name, c in counts.items():
    print("%-58s %6d " % (name.replace("b200::", "")[:58], c["_n"]) + " ".join("%7d" % c[k] for k in KEYS))
print("%-58s %6s " % ("TOTAL", "") + " ".join("%7d" % total[k] for k in KEYS))

"""Per-kernel SASS opcode histogram of libgd_b200.so (cuobjdump -sass): which kernels are Blackwell-native
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA) and which use the legacy warp-level
tensor path (HMMA = mma.sync).  Usage: python profiles/sass_histogram.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "guided_diffusion_clip_b200", "libgd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS",
        "SYNCS", "USETMAXREG"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            hist[kern]["UTCHMMA.2CTA"] += 1
        total[kern] += 1
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>12s}" for k in KEYS))
for k, h in hist.items():
    print(f"{k[:70]:70s} {total[k]:7d} " + " ".join(f"{h.get(x, 0):12d}" for x in KEYS))

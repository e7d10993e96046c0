#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libmdconv_b200.so (cuobjdump -sass), for profiles/: the Blackwell-native
evidence (UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,
USETMAXREG = setmaxnreg, SYNCS = mbarrier) next to the ordinary memory opcodes.  Runs on the CPU box.
usage: python scripts/sass_counts.py [lib] > profiles/r2_sass_opcounts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "modulated-deform-conv_b200", "lib", "libmdconv_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "USETMAXREG",
         "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDG", "RED", "ATOMG", "ATOMS", "ATOM", "HFMA2", "FFMA", "MEMBAR",
         "FENCE", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kern = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        counts[kern][op] += 1
        counts[kern]["_total"] += 1
names = list(counts)
try:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
except Exception:
    demangle = {n: n for n in names}


def short(n):
    s = demangle.get(n, n)
    s = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", s)
    s = re.sub(r"^void ", "", s)
    s = s.replace("mdconv::", "").replace("tc::", "")
    m = re.match(r"([A-Za-z_0-9]+)(<.*?>)?\(", s)
    if m:
        args = (m.group(2) or "").replace("(int)", "").replace("(bool)", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16")
        return m.group(1) + args.replace(" ", "")
    return s


print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: opcode counts per kernel (static instruction counts)")
tot = collections.Counter()
rows = []
for n in names:
    c = counts[n]
    for k, v in c.items():
        tot[k] += v
    rows.append((short(n), c))
rows.sort(key=lambda r: r[0])
cols = [w for w in WATCH if tot[w] > 0]
print("kernel".ljust(64) + " " + " ".join(c.rjust(8) for c in ["instrs"] + cols))
for name, c in rows:
    if not name.startswith("k_"):
        continue
    print(name[:64].ljust(64) + " " + " ".join(str(c[k]).rjust(8) for k in ["_total"] + cols))
print("TOTAL (all functions incl. cub)".ljust(64) + " " + " ".join(str(tot[k]).rjust(8) for k in ["_total"] + cols))

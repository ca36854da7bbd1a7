"""CPU: count the Blackwell-native SASS mnemonics per kernel of libdtraj.so (cuobjdump -sass): tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, cp.async -> LDGSTS, legacy tensor path -> HMMA (must be absent).
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distillation_trajectories_b200 import _lib

out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "LDGSTS", "SYNCS", "HMMA", "total"]
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("dtraj::", "").replace("void ", "")
        per[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in ln:
        continue
    op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not op:
        continue
    o = op.group(1)
    per[cur]["total"] += 1
    for k in keys[:-1]:
        if k == "UTCHMMA.2CTA":
            if o.startswith("UTCHMMA") and ".2CTA" in o:
                per[cur][k] += 1
        elif k == "HMMA":
            if o.startswith("HMMA"):
                per[cur][k] += 1
        elif o.startswith(k):
            per[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(_lib.LIB_PATH)} (sm_100a): instruction counts per kernel")
print(f"{'kernel':44s} " + " ".join(f"{k:>12s}" for k in keys))
tot = collections.Counter()
for name, c in per.items():
    print(f"{name[:44]:44s} " + " ".join(f"{c[k]:12d}" for k in keys))
    tot.update(c)
print(f"{'ALL':44s} " + " ".join(f"{tot[k]:12d}" for k in keys))

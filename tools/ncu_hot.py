"""Read a .ncu-rep (source page) and print (a) totals, (b) sample-heavy SASS regions grouped by execution count.
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [min_samples]"""
import csv
import subprocess
import sys

path = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
tot = sum(int(r[isamp] or 0) for r in data)
totex = sum(int(r[iex] or 0) for r in data)
print("kernel", rows[0][1][:80], "samples", tot, "warp-instr", totex, "sass", len(data))
cur, acc, n, start, texts = None, 0, 0, None, []
regions = []
for r in data:
    ex, s = int(r[iex] or 0), int(r[isamp] or 0)
    if ex != cur:
        if cur is not None:
            regions.append((start, cur, n, acc, texts))
        cur, acc, n, start, texts = ex, 0, 0, r[ia][-5:], []
    acc += s
    n += 1
    texts.append((s, r[isrc]))
regions.append((start, cur, n, acc, texts))
for st, ex, n, a, texts in regions:
    if a >= thr:
        top = sorted(texts, key=lambda t: -t[0])[:3]
        print(f"{st} exec {ex:9d} n {n:4d} samples {a:6d} ({a / tot * 100:4.1f}%)  " + " | ".join(f"{s}:{t.strip()[:40]}" for s, t in top))

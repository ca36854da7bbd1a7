"""Average the runs of a tools/ab.sh log per (model, layer) and print A, B and the relative change.
    bash tools/ab.sh libA.so libB.so "conv|enc1" 16 592 1.0,0.5 > log.txt; python tools/ab_table.py log.txt"""
import collections
import re
import sys

cur = model = None
data = collections.defaultdict(lambda: collections.defaultdict(list))
order = []
for l in open(sys.argv[1]):
    if l.startswith("##"):
        cur = l.split()[1]
        if cur not in order:
            order.append(cur)
    elif l.startswith("=="):
        model = l.split()[1]
        data[cur][(model, "TOTAL forward + step")].append(float(l.split(":")[1].split()[0]))
    else:
        m = re.match(r"^(\S.*?)\s+(\d+)\s+([\d.]+)\s+(\d+)\s+[\d.]+%", l)
        if m:
            data[cur][(model, m.group(1).strip())].append(float(m.group(3)))
a_name, b_name = order[0], order[1]
print(f"# A = {a_name}, B = {b_name}: microseconds, mean of {len(next(iter(data[a_name].values())))} alternating runs on one GPU box")
for k in data[a_name]:
    a, b = data[a_name][k], data[b_name][k]
    ma, mb = sum(a) / len(a), sum(b) / len(b)
    print(f"{k[0]:8s} {k[1]:28s} A {ma:8.1f}  B {mb:8.1f}  {100 * mb / ma - 100:+5.1f}%")

"""Table of a knob sweep: tools/ab_table2.py <file> [layer regex]  -- the file holds blocks '## <setting> rep <k>' followed by
tools/profile_layers.py output; prints, per (model, layer), the microseconds of every setting (repetitions joined by '/')."""
import collections, re, sys
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
cur = model = None
data = collections.OrderedDict()
cfgs = []
for l in open(sys.argv[1]):
    if l.startswith("##"):
        cur = l[2:].rsplit("rep", 1)[0].strip()
        if cur not in cfgs: cfgs.append(cur)
        continue
    if l.startswith("=="):
        model = l.split()[1]
        tot = re.search(r": (\d+) us", l)
        data.setdefault((model, "TOTAL"), collections.defaultdict(list))[cur].append(float(tot.group(1)))
        continue
    m = re.match(r"(\S.*?)\s{2,}(\d+)\s+([\d.]+)", l)
    if m and (pat is None or pat.search(m.group(1))):
        data.setdefault((model, m.group(1)), collections.defaultdict(list))[cur].append(float(m.group(3)))
w = max(len(c) for c in cfgs) + 1
print("layer".ljust(34), "".join(c.ljust(w) for c in cfgs))
for k, v in data.items():
    print((k[0] + " " + k[1]).ljust(34), "".join("/".join("%.0f" % x for x in v.get(c, [])).ljust(w) for c in cfgs))

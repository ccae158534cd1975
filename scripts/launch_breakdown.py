import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = OrderedDict()
for r in rows[hdr + 1:]:
    name = r[4].split("(")[0]
    agg.setdefault(name, []).append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-40s n=%3d mean %10.1f us  share %.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))

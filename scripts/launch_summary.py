"""per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py file.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) > vi:
        try:
            agg[r[ki][:60]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in agg.values())
print("launches", sum(len(v) for v in agg.values()), "total ms %.3f" % (tot / 1e6))
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1]))[:top]:
    print("%-62s n=%5d mean %9.2f us total %8.3f ms %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e6, 100 * sum(v) / tot))

"""Per-kernel summary of an ncu --csv launch list (any number of metrics per launch)."""
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rr = csv.reader(lines); h = next(rr)
ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rr:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    d = per.setdefault((r[idi], name), {})
    d[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), d in per.items():
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    for k, v in d.items():
        a[k] += v
for name, a in agg.items():
    n = a["n"]
    print("%-44s n=%3d  avg %9.1f us  rd %8.1f MB  wr %8.1f MB" % (name, n, a.get("gpu__time_duration.sum", 0) / n / 1e3,
          a.get("dram__bytes_read.sum", 0) / n / 1e6, a.get("dram__bytes_write.sum", 0) / n / 1e6))

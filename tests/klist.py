import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rr = csv.reader(lines); h = next(rr)
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
rows = [(r[ki].split("(")[0], float(r[vi].replace(",", ""))) for r in rr if len(r) > vi]
seen = {}
for n, v in rows:
    seen.setdefault(n, []).append(v)
for n, v in seen.items():
    print("%-44s n=%3d last %9.1f us  max %9.1f" % (n, len(v), v[-1] / 1e3, max(v) / 1e3))

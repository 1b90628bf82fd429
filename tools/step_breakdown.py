"""Print one training step's kernels from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
recs = [(r["Kernel Name"], r["Block Size"], r["Grid Size"], float(r["Metric Value"])) for r in csv.DictReader(lines)
        if r.get("Metric Name") == "gpu__time_duration.sum"]
ad = [i for i, x in enumerate(recs) if "grad_reduce_adam" in x[0]]
s, e = ad[0] + 1, ad[1] + 1
tot = 0.0
agg = {}
for name, blk, grid, ns in recs[s:e]:
    short = re.sub(r"\(.*", "", name).replace("s2s::", "").replace("void ", "").replace("<unnamed>::", "")
    if "-v" in sys.argv:
        print(f"{ns / 1000:8.2f} us  grid={grid:>14} blk={blk:>12}  {short[:70]}")
    tot += ns
    k = short.split("<")[0]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += ns / 1000
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} n={n:3d} total={us:8.1f} us  avg={us / n:6.2f} us")
print("sum of kernel durations (us):", round(tot / 1000, 1), "launches:", e - s)

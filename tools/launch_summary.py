"""Aggregate an ncu launch list (gpu__time_duration.sum csv) per kernel; optionally list launches over a threshold.
    python tools/launch_summary.py gpurun_out/x.csv [min_ms] [--second-half]"""
import collections, csv, sys
path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else None
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
if "--second-half" in sys.argv:
    rows = rows[len(rows) // 2:]
agg = collections.defaultdict(lambda: [0, 0.0])
for x in rows:
    v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    x["ms"] = v
    n = x["Kernel Name"].split("(")[0][:64]
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-66s %4d %9.3f ms %5.1f%%" % (n, c, t, 100 * t / tot))
print("total %.3f ms over %d launches" % (tot, len(rows)))
if thr is not None:
    for x in rows:
        if x["ms"] > thr:
            print("%-50s %-16s %8.3f" % (x["Kernel Name"][:50], x["Grid Size"], x["ms"]))

"""Per-kernel table from an ncu multi-metric launch list (csv, one row per launch x metric):
    python tools/kernel_metrics.py gpurun_out/x.csv [hbm_peak_gbs]
Aggregates by kernel: launches, total ms, DRAM GB moved, achieved DRAM GB/s (and % of the measured peak), tensor pipe %."""
import collections, csv, json, os, sys
second_half = "--second-half" in sys.argv   # two identical steps were profiled: keep the second (no first-call setup)
sys.argv = [a for a in sys.argv if a != "--second-half"]
path = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6457.4
with open(path) as f:
    rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]; n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
    if n.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[n] = v
if second_half:
    keys = list(launch.keys())
    for k in keys[:len(keys) // 2]:
        del launch[k]
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"][:58], {"n": 0, "ms": 0.0, "bytes": 0.0, "tc": 0.0})
    ms = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1; a["ms"] += ms
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["tc"] += ms * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                          d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0))
tot = sum(a["ms"] for a in agg.values())
print("%-58s %4s %9s %6s %9s %9s %6s %7s" % ("kernel", "n", "ms", "%step", "DRAM GB", "GB/s", "%HBM", "tensor%"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = a["bytes"] / 1e9 / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
    print("%-58s %4d %9.3f %6.1f %9.2f %9.0f %6.1f %7.1f" % (n, a["n"], a["ms"], 100 * a["ms"] / tot, a["bytes"] / 1e9, gbs,
                                                            100 * gbs / peak, a["tc"] / a["ms"] if a["ms"] else 0.0))
tb = sum(a["bytes"] for a in agg.values())
print("total %.3f ms, %d launches, DRAM %.1f GB -> %.0f GB/s = %.1f %% of the measured %.0f GB/s" %
      (tot, len(launch), tb / 1e9, tb / 1e9 / (tot / 1e3), 100 * tb / 1e9 / (tot / 1e3) / peak, peak))

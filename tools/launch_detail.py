"""Per-launch rows (second half = second profiled step) of an ncu multi-metric launch list, filtered by substrings.
    python tools/launch_detail.py gpurun_out/x.csv [substr ...]"""
import collections, csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
pats = sys.argv[2:]
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0][:52], "grid": r.get("Grid Size", "")})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]; n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
    if n.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[n] = v
keys = list(L.keys())
for k in keys[len(keys) // 2:]:
    d = L[k]
    if pats and not any(p in d["name"] for p in pats):
        continue
    print("%5s %-52s %-14s %7.3f ms  rd %6.2f GB wr %6.2f GB  tensor %3.0f%%" % (
        k, d["name"], d["grid"], d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0) / 1e9,
        d.get("dram__bytes_write.sum", 0) / 1e9, d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)))

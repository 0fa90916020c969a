"""Summarise an ncu report here (no GPU): key metrics + the top stall sites of the SASS/source page.
    python tools/ncu_top.py gpurun_out/x.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, u, v = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__cycles_active.avg", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for i, n in enumerate(h):
    if n in want:
        print("%-90s %-8s %s" % (n, u[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(src.splitlines()))
hi = next(i for i, x in enumerate(r) if "Source" in x and "# Samples" in x)
h = r[hi]; rows = [x for x in r[hi + 1:] if len(x) == len(h)]
k, s = h.index("# Samples"), h.index("Source")
tot = sum(float(x[k] or 0) for x in rows)
print("total samples", tot, "sass rows", len(rows))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(float(x[h.index(n)] or 0) for x in rows) for n in stalls}
print("stall totals:", {n: int(a) for n, a in sorted(agg.items(), key=lambda kv: -kv[1]) if a > 0.01 * tot})
order = sorted(range(len(rows)), key=lambda i: -float(rows[i][k] or 0))
for i in order[:topn]:
    x = rows[i]
    st = {n[6:]: x[h.index(n)] for n in stalls if x[h.index(n)] not in ("0", "")}
    print("%5d %7s  %-70s %s" % (i, x[k], x[s][:70], st))

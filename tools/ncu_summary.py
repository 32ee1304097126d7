"""Summarise an .ncu-rep (read on the CPU box): per-kernel key metrics + stall breakdown.  usage: ncu_summary.py rep [out.json]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for r in data:
    d = {"kernel": r[h.index("Kernel Name")][:90]}
    for k in keys:
        if k in h:
            d[k] = r[h.index(k)] + " " + units[h.index(k)]
    st = {}
    for i, n in enumerate(h):
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
            st[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(r[i] or 0), 2)
    d["stalls_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:7])
    out.append(d)
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)

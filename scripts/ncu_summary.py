"""Markdown summary of one kernel from an .ncu-rep (raw page): python scripts/ncu_summary.py <rep> <title>"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.max.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
print(f"# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none --import-source on`, not tracked; summary below).\n")
print("| metric | value |\n|---|---|")
d = dict(zip(hdr, zip(units, vals)))
for w in want:
    if w in d:
        print(f"| {w} | {d[w][1]} {d[w][0]} |")
print("\nWarp stall reasons (warps stalled per issue-active cycle, > 0.25):\n\n| reason | ratio |\n|---|---|")
st = []
for h, (u, v) in d.items():
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
        try:
            st.append((float(v), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
for v, n in sorted(st, reverse=True):
    if v > 0.25:
        print(f"| {n} | {v:.2f} |")

"""Summary of one kernel from an .ncu-rep (raw page):

    python scripts/ncu_summary.py <rep> <title> [--json profiles/ncu_r02.json --key <workload>:<kernel> --command "<cmd>"]

Prints a markdown table (redirect it into profiles/<name>.md) and, with --json, records the facts bench.py quotes
(DRAM bytes per launch, pipe utilisation) under `key` in the tracked JSON together with the capture's command, report
name and the commit it was taken at -- bench.py reads them from there instead of carrying literals."""
import csv, json, os, subprocess, sys

args = sys.argv[1:]
rep, title = args[0], args[1]
opt = {args[i]: args[i + 1] for i in range(2, len(args) - 1, 2)}
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.max.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
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


def num(name, scale=1.0):
    try:
        u, v = d[name]
        v = float(v.replace(",", ""))
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return v * mult * scale
    except Exception:
        return None


if "--json" in opt:
    path = opt["--json"]
    db = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    db[opt["--key"]] = {
        "report": os.path.basename(rep), "command": opt.get("--command", ""), "commit": commit,
        "kernel_name": d.get("Kernel Name", ("", ""))[1][:80], "duration_us_under_ncu": num("gpu__time_duration.sum"),
        "dram_bytes_per_launch": (rd or 0) + (wr or 0) if rd is not None else None,
        "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lts_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active")}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)

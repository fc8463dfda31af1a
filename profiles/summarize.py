#!/usr/bin/env python
"""Turns ncu outputs under gpurun_out/ into the small text summaries kept in profiles/.
   python profiles/summarize.py <tag>"""
import collections, csv, subprocess, sys, os

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(root, "gpurun_out")
P = os.path.join(root, "profiles")

def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"].split("(")[0]
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, cold cache, serialised): compare SHARES\n")
        f.write("# source: %s\n" % os.path.basename(path))
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write("%-28s launches=%5d total_ms=%10.3f share=%5.1f%%\n" % (k, cnt[k], v, 100 * v / T))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

def kernel(rep, out):
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(res.splitlines()))
    if len(r) < 3:
        return
    hdr = r[0]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % os.path.basename(rep))
        for row in r[2:]:
            f.write("kernel: %s\n" % row[hdr.index("Kernel Name")][:110])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("  %-84s %-16s %s\n" % (w, r[1][i], row[i]))

for name in os.listdir(G):
    if name.startswith("launches_") and name.endswith(tag + ".csv"):
        launches(os.path.join(G, name), os.path.join(P, name.replace(".csv", ".summary.txt")))
    if name.endswith(tag + ".ncu-rep"):
        kernel(os.path.join(G, name), os.path.join(P, name.replace(".ncu-rep", ".summary.txt")))
print(sorted(os.listdir(P)))

"""Print key metrics + per-instruction stall hot spots from an .ncu-rep (build container)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ksel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []   # one kernel of a multi-kernel report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
for k in ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_elapsed",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
          "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "launch__grid_size"]:
    print(k, d.get(k), u.get(k))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))   # first (SASS) block only
data = [r for r in rows[2:end] if len(r) >= len(hdr) - 2]; ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
mx = max(int(r[ix["Instructions Executed"]]) for r in data)
loop = [r for r in data if int(r[ix["Instructions Executed"]]) >= 0.95 * mx]
print("total samples", tot, "loop instrs", len(loop), "samples in loop", sum(int(r[ix["# Samples"]]) for r in loop))
def op(r):
    t = r[ix["Source"]].split()
    if t[0].startswith("@"): t = t[1:]
    return t[0].split(".")[0]
print(collections.Counter(op(r) for r in loop).most_common(14))
agg = collections.Counter()
for r in data:
    for k in ("stall_wait", "stall_math", "stall_selected", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_not_selected", "stall_no_inst", "stall_dispatch", "stall_mio", "stall_lg", "stall_misc", "stall_sleep", "stall_barrier"):
        if k in ix: agg[k] += int(r[ix[k]])
print(agg.most_common())
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    print(r[ix["# Samples"]], r[ix["Source"]].strip()[:60], "| wait", r[ix["stall_wait"]], "math", r[ix["stall_math"]], "ssb", r[ix["stall_short_sb"]], "lsb", r[ix["stall_long_sb"]], "br", r[ix["stall_branch_resolving"]], "nosel", r[ix["stall_not_selected"]])

#!/usr/bin/env python
"""Summarise gpurun_out/ ncu captures into profiles/ (run in the build container).

    python tools/summarize_ncu.py <tag> <prof.ncu-rep> <launches.csv> <B> <T> <m> <k>
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.min.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    tag, rep, launches = sys.argv[1:4]
    wl = [int(v) for v in sys.argv[4:8]]
    os.makedirs("profiles", exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    traffic = None
    total_traffic = 0.0
    pipe_weighted = 0.0
    dur_total = 0.0
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name"), "id": d.get("ID")}
        for k in KEYS:
            if k in d:
                rec[k] = d[k] + " " + u[k]
        out.append(rec)

        def gb(key):
            v, un = float(d[key].replace(",", "")), u[key]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[un]

        traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        total_traffic += traffic
        try:
            dur = float(d["gpu__time_duration.sum"].replace(",", ""))
            dur_total += dur
            pipe_weighted += dur * float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"])
        except Exception:
            pass
    with open("profiles/%s_ncu_full_summary.json" % tag, "w") as fh:
        json.dump({"source": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on "
                   "-k regex:dfm_ -s 9 -c 3 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-other",
                   "workload_B_T_m_k": wl, "launches": out}, fh, indent=1)
    with open("profiles/ncu_traffic.json", "w") as fh:
        pipe = pipe_weighted / dur_total if dur_total > 0 else None
        # one step of the hot path = all captured launches (transient + steady + resume kernels)
        json.dump({"workload": wl, "dram_bytes_per_launch": total_traffic, "fp64_pipe_active_pct": pipe,
                   "launches_summed": len(out),
                   "from": "profiles/%s_ncu_full_summary.json" % tag}, fh)
    # launch list: per-kernel totals and shares
    tot = {}
    n = 0
    with open(launches) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r.get("Metric Unit", "ns"), 1e-6)
        name = r["Kernel Name"].split("(")[0]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
        n += 1
    total = sum(v[1] for v in tot.values())
    with open("profiles/%s_launches_summary.md" % tag, "w") as fh:
        fh.write("# %s: ncu launch list of `python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu`\n\n" % tag)
        fh.write("(`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400`; "
                 "bench.py brackets warm-up + timed steps with cudaProfilerStart/Stop so the torch kernels that "
                 "generate the synthetic y are excluded; cold-cache, serialised: compare shares, not absolutes)\n\n")
        fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
            fh.write("| `%s` | %d | %.3f | %.1f %% |\n" % (name[:90], c, ms, 100 * ms / total))
        fh.write("\n%d launches captured, %.2f ms total.\n" % (n, total))
    print("traffic", traffic, "launches", n)


if __name__ == "__main__":
    main()

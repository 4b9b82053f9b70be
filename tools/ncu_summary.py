"""Condense ncu reports into one CSV row per captured launch (the files under profiles/ are made with this).

    python tools/ncu_summary.py out.csv report1.ncu-rep [report2.ncu-rep ...]
"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

out, reps = sys.argv[1], sys.argv[2:]
rows = []
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, units = r[0], r[1]
    for row in r[2:]:
        d = {"report": rep.split("/")[-1], "kernel": row[h.index("Kernel Name")][:120]}
        for m in METRICS:
            if m in h:
                d[m] = (row[h.index(m)] + " " + units[h.index(m)]).strip()
        rows.append(d)
with open(out, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=["report", "kernel"] + METRICS)
    w.writeheader()
    w.writerows(rows)
print(f"{len(rows)} launches -> {out}")

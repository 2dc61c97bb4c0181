#!/usr/bin/env python
"""Text summary (one block per profiled launch) of an `ncu --set full` report, for profiles/.
usage: tools/ncu_kernels_summary.py report.ncu-rep out.txt ["label for launch 0" "label for launch 1" ...]"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
labels = sys.argv[3:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
lines = [f"# ncu --set full --clock-control none summary of {rep} (one profiled launch per block; cold-cache, serialised replays)"]
for i, r in enumerate(rows[2:]):
    lines.append(f"launch {i}: {labels[i] if i < len(labels) else ''}")
    lines.append("  kernel: " + r[hdr.index("Kernel Name")][:150])
    for k in KEYS:
        if k in hdr:
            lines.append(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

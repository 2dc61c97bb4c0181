#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into a small text file for profiles/.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt [launches.csv]
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep} (per launch; cold-cache, serialised replays)"]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:160])
        for k in KEYS:
            if k in hdr:
                lines.append(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
        if "dram__bytes_read.sum" in hdr:
            def val(k):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            lines.append(f"  traffic(dram read+write) = {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):.0f} bytes")
    if len(sys.argv) > 3:
        lines.append(f"# launch list {sys.argv[3]} (ncu --metrics gpu__time_duration.sum --clock-control none)")
        agg = collections.defaultdict(list)
        h = None
        for r in csv.reader(open(sys.argv[3])):
            if r and r[0] == "ID":
                h = r
                continue
            if h and len(r) == len(h):
                d = dict(zip(h, r))
                agg[(d["Kernel Name"][:110], d["Grid Size"], d["Block Size"])].append(float(d["Metric Value"]))
        tot = sum(sum(v) for v in agg.values())
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"  {sum(v) / tot * 100:5.1f}%  n={len(v):4d} avg={sum(v) / len(v) / 1e3:9.2f} us  {k[0]} grid={k[1]} block={k[2]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

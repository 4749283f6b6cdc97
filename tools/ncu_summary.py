#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the few counters DESIGN.md cites.
usage: tools/ncu_summary.py report.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append(f"== {r[H.index('Kernel Name')][:110]}")
        for k in KEYS:
            if k in H:
                i = H.index(k)
                out.append(f"  {k:92s} {r[i]:>18s} {U[i]}")
        for i, h in enumerate(H):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("0", "0.000000"):
                out.append(f"  {h:92s} {r[i]:>18s} {U[i]}")
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(f"# ncu --set full --clock-control none, summary of {rep.split('/')[-1]} (tools/ncu_summary.py)\n" + txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()

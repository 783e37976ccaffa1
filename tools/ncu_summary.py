#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) into the handful of
counters DESIGN.md argues from.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
    ("smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "global load efficiency %"),
    ("smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct", "global store efficiency %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary of %s (per launch; cold-cache, serialised replay)" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print("\n== %s  grid %s block %s" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-32s %s %s" % (label, r[i], units[i]))
        stalls = []
        for i, k in enumerate(hdr):
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stall reasons (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))


if __name__ == "__main__":
    main()

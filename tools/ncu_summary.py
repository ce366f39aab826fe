#!/usr/bin/env python
"""tools/ncu_summary.py -- text summary of an `ncu --set full` report for profiles/.

    ncu --set full --clock-control none --import-source on \
        --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum \
        -k regex:<kernel> -c <n> -o gpurun_out/<name> <command>            (on the GPU box, under gpurun)
    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [kernel-substring] > profiles/rNN_<name>_ncu_full.txt   (here)

One block per profiled launch: `kernel: <name>  grid (...) block (...)`, then `  <metric>  <unit>  <value>` lines (the format
bench.py's profile_metrics() parses) and the warp stall reasons per issue-active cycle, largest first.
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    print("# from %s (ncu --set full --clock-control none)" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if want and want not in name:
            continue
        print("kernel: %s  grid %s block %s" % (name, r[col["Grid Size"]], r[col["Block Size"]]))
        for m in METRICS:
            if m in col and r[col[m]] != "":
                print("  %-75s %-12s %s" % (m, units[col[m]], r[col[m]]))
        st = []
        for n, i in col.items():
            if n.startswith(STALL) and n.endswith("_per_issue_active.ratio") and r[i] not in ("", "0"):
                st.append((float(r[i].replace(",", "")), n[len(STALL):-len("_per_issue_active.ratio")]))
        st.sort(reverse=True)
        print("  warp stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))


if __name__ == "__main__":
    main()

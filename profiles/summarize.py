#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_train_r1.ncu-rep > profiles/r01_train_kernel.txt

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed) and prints, per captured
launch, the metrics DESIGN.md and bench.py's roofline refer to.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def main():
    path = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    index = {name: k for k, name in enumerate(header)}
    print(f"# ncu --set full --clock-control none summary of {path}")
    for row in rows[2:]:
        print(f"\n## launch {row[index['ID']]}: {row[index['Kernel Name']]}")
        for key in KEYS:
            if key in index:
                print(f"{key:90s} {row[index[key]]:>20s} {units[index[key]]}")
        read = float(row[index["dram__bytes_read.sum"]])
        write = float(row[index["dram__bytes_write.sum"]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        traffic = read * scale[units[index["dram__bytes_read.sum"]]] + \
            write * scale[units[index["dram__bytes_write.sum"]]]
        print(f"{'traffic = dram read + write per launch':90s} {traffic:20.0f} byte")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel table of the metrics that matter here from an .ncu-rep file (ncu --set full): duration, DRAM bytes,
pipe utilisation, hit rates, occupancy, stall reasons.  python scripts/ncu_table.py report.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_requests_srcunit_tex.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]


def main():
    report = sys.argv[1]
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    index = {h: i for i, h in enumerate(header)}
    print("# %s (ncu --set full --clock-control none; per launch)" % report.split("/")[-1])
    for row in rows[2:]:
        print("== %s   grid %s" % (row[index["Kernel Name"]][:110], row[index["launch__grid_size"]]))
        for name in WANT:
            if name in index:
                print("   %-86s %s %s" % (name, row[index[name]], units[index[name]]))


if __name__ == "__main__":
    main()

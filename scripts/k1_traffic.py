#!/usr/bin/env python
"""profiles/k1_walk_traffic.json from an ncu --set full report that holds one launch of the walk kernel on config 2:
DRAM bytes per launch (bench.py's roofline.traffic) and the L2 hit rate of its requests.
  python scripts/k1_traffic.py gpurun_out/r02_step_kernels.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    report = sys.argv[1]
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    index = {h: i for i, h in enumerate(header)}

    def value(row, name, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "%": 1.0}):
        return float(row[index[name]]) * scale.get(units[index[name]], 1.0)
    for row in rows[2:]:
        name = row[index["Kernel Name"]]
        if "k1_walk" not in name:
            continue
        read, write = value(row, "dram__bytes_read.sum"), value(row, "dram__bytes_write.sum")
        out = {"kernel": name.split("(")[0].replace("void ", ""),
               "workload": "config 2 (|B| = 1.51e9 inserted bases), one launch",
               "dram_bytes_read": int(read), "dram_bytes_write": int(write), "dram_bytes_per_launch": int(read + write),
               "algorithmic_bytes_per_launch": int(168.0 * 1510000000),
               "l2_hit_rate": value(row, "lts__t_sector_hit_rate.pct") / 100.0,
               "gpu_time_ms": value(row, "gpu__time_duration.sum"),
               "registers_per_thread": int(float(row[index["launch__registers_per_thread"]])),
               "source": "%s (ncu --set full --clock-control none --import-source on)" % os.path.basename(report)}
        json.dump(out, open(os.path.join(ROOT, "profiles", "k1_walk_traffic.json"), "w"), indent=1)
        print(json.dumps(out, indent=1))
        return
    raise SystemExit("no walk kernel in " + report)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of per-launch numbers DESIGN.md / bench.py cite.

usage: python tools/ncu_summary.py gpurun_out/tc_full.ncu-rep profiles/r01/ncu_conv_tc.json
Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps, per launch: duration, DRAM bytes read /
written, registers, grid/block, SM / tensor-pipe / L2 / DRAM utilisation."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "sm__cycles_active.avg": "sm_cycles_active",
    "smsp__inst_executed.sum": "warp_insts",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0,
         "nsecond": 1e-3}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0]}
        for k, name in KEEP.items():
            if k in col and r[col[k]] != "":
                try:
                    v = float(r[col[k]].replace(",", ""))
                except ValueError:
                    continue
                d[name] = v * SCALE.get(units[col[k]], 1.0)
        if "dram_read" in d and "dram_write" in d:
            d["traffic_bytes"] = d["dram_read"] + d["dram_write"]
        launches.append(d)
    json.dump({"source": rep, "how": "ncu --set full --clock-control none --import-source on; ncu -i --page raw --csv",
               "launches": launches}, open(out, "w"), indent=1)
    for d in launches:
        print(d)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

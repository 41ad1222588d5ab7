#!/usr/bin/env python
"""Condense an ncu report into the handful of counters DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_xyz.json

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per profiled launch, the
duration, DRAM bytes, tensor-pipe / memory-pipe activity, registers and occupancy figures.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    result = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")],
               "block": r[hdr.index("Block Size")]}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                rec[h] = {"value": v, "unit": u}
        result.append(rec)
    with open(out, "w") as f:
        json.dump({"source": rep, "launches": result}, f, indent=1)
    for rec in result:
        print(rec["kernel"][:60], {k: v["value"] for k, v in rec.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()

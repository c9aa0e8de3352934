#!/usr/bin/env python
"""profiles/traffic.json from `ncu --page raw --csv` exports of the fused lnpost kernel (development tool):

    python tools/update_traffic.py posterior_like=profiles/r2x_lnpost_posterior_raw.csv grid_wide=profiles/r2x_lnpost_grid_wide_raw.csv

Records, per batch, the DRAM bytes of one launch (dram__bytes_read.sum + dram__bytes_write.sum) and the counters the
roofline block of bench.py quotes, together with a hash of the CUDA sources the capture was made from: bench.py reports
`traffic` only while that hash matches the sources it is timing."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

COUNTERS = {
    "duration_us": "gpu__time_duration.sum",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1_wavefront_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_instructions": "smsp__inst_executed.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "l2_sectors_read_from_l1": "lts__t_sectors_srcunit_tex_op_read.sum",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {k: (u, v) for k, u, v in zip(hdr, units, vals)}


def main():
    out = {"kernel_source_hash": bench.kernel_source_hash(), "batches": {}}
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        d = read(path)
        dram = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u, v = d[k]
            dram += float(v) * UNIT_SCALE[u]
        counters = {}
        for label, key in COUNTERS.items():
            if key in d:
                try:
                    counters[label] = float(d[key][1])
                except ValueError:
                    pass
        out["batches"][name] = {"dram_bytes_per_launch": int(round(dram)), "kernel": d["Kernel Name"][1].strip(),
                                "capture": os.path.relpath(path, ROOT), "rows_per_launch": bench.BATCH,
                                "algorithmic_bytes_per_launch": int(bench.B_ALG * bench.BATCH), "counters": counters}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

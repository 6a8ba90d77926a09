#!/usr/bin/env python
"""Condenses an `ncu --set full` report of one pass-kernel launch into the JSON the repository
keeps under profiles/ (and bench.py reads for its issue-slot ceiling).  Needs only the `ncu`
command-line tool, no GPU:

    python tools/ncu_summarise.py gpurun_out/ncu/f32_pass.ncu-rep --pixels 16777216 --taps 65 \
        > profiles/r2_pass_kernel_ncu_summary.json
    python tools/ncu_summarise.py REPORT --details > profiles/rN_..._ncu_details.txt
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_rows(report: str) -> list[dict]:
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True)
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units = rows[0], rows[1]
    return [dict(zip(header, r)) for r in rows[2:]], dict(zip(header, units))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--pixels", type=int, default=0, help="pixels the launch computed")
    ap.add_argument("--taps", type=int, default=0)
    ap.add_argument("--details", action="store_true", help="print ncu's details page instead")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    if args.details:
        sys.stdout.write(subprocess.run(["ncu", "-i", args.report, "--page", "details"], capture_output=True,
                                        text=True, check=True).stdout)
        return
    rows, units = raw_rows(args.report)
    row = rows[0]
    summary = {"kernel": row.get("Kernel Name", ""), "report": args.report, "note": args.note}
    for key in KEEP:
        if key in row and row[key] != "":
            unit = units.get(key, "")
            value = row[key].replace(",", "")
            try:
                number = float(value)
            except ValueError:
                summary[key] = f"{row[key]} {unit}".strip()
                continue
            if key.startswith("dram__bytes") or key == "lts__t_bytes.sum":
                summary[key] = number * BYTES.get(unit, 1)
            else:
                summary[key] = f"{number:g} {unit}".strip()
    stalls = {}
    for key, value in row.items():
        if key.startswith(STALL) and key.endswith("_per_issue_active.ratio") and value not in ("", "n/a"):
            name = key[len(STALL):-len("_per_issue_active.ratio")]
            try:
                stalls[name] = round(float(value.replace(",", "")), 3)
            except ValueError:
                pass
    summary["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
    if args.pixels and args.taps and "smsp__inst_executed.sum" in row:
        inst = float(row["smsp__inst_executed.sum"].replace(",", ""))
        summary["warp_instructions_per_pixel_step"] = inst / (args.pixels * (args.taps - 1) / 32)
        summary["pixels"], summary["taps"] = args.pixels, args.taps
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()

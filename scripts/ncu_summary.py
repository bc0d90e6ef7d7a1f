#!/usr/bin/env python
"""Compact JSON summary of an `ncu --set full` report: one record per captured launch with the metrics DESIGN.md cites.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN/<name>_ncu_full_summary.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                rec[w] = r[i] if w == "Kernel Name" else f"{r[i]} {units[i]}".strip()
        if "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum" in rec:
            try:
                req = float(rec["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"].split()[0].replace(",", ""))
                sec = float(rec["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"].split()[0].replace(",", ""))
                rec["sectors_per_request_global_ld"] = round(sec / req, 3) if req else None
            except Exception:
                pass
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()

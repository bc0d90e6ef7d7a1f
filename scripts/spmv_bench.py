#!/usr/bin/env python
"""Kernel-level A/B bench of the SpMV kernel modes on the BASELINE workload families (GPU box only).

    python scripts/spmv_bench.py [--modes default,perblock,window] [--workloads C2:1,C3:5,C5:20,C4:10] [--reps 20]

For every workload (name:scale) and mode (environment switches read at initialize) it times y += A x (mode 1) and x += A'y (mode 2) through the
C ABI (lsqr_b200_ez_aprod on device vectors) with CUDA events on the stream the kernels run on, reports
algorithmic GB/s against MEASURED_PEAKS.json, cross-checks the modes against each other, and times a
full solve.  One JSON line per (workload, mode).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import lsqr_b200
from lsqr_b200 import synth, synth_device


MODES = {
    "default": {},
    "perblock": {"LSQR_B200_SINGLE_LAUNCH": "0"},
    "noguard": {"LSQR_B200_DRIFT_GUARD": "0"},
    "guard2": {"LSQR_B200_DRIFT_GUARD": "2"},
    "nooverlap": {"LSQR_B200_OVERLAP_UPDATE": "0"},
    "window": {"LSQR_B200_WINDOW": "1"},
    "pdl": {"LSQR_B200_PDL": "1"},
    "local": {"LSQR_B200_FLAVOUR": "0"},
    "batch1": {"LSQR_B200_BATCH": "1"},
    "batch2": {"LSQR_B200_BATCH": "2"},
    "batch5": {"LSQR_B200_BATCH": "5"},
    "batch8": {"LSQR_B200_BATCH": "8"},
    "batch12": {"LSQR_B200_BATCH": "12"},
    "tile4k": {"LSQR_B200_WARP_TILE": "4096"},
    "tile16k": {"LSQR_B200_WARP_TILE": "16384"},
    "tile32k": {"LSQR_B200_WARP_TILE": "32768"},
    "roww2": {"LSQR_B200_TILE_ROW_WEIGHT": "2"},
    "roww8": {"LSQR_B200_TILE_ROW_WEIGHT": "8"},
    "gather": {"LSQR_B200_FLAVOUR": "2"},
    "carve28": {"LSQR_B200_SMEM_CARVEOUT_PCT": "28"},
    "carve24": {"LSQR_B200_SMEM_CARVEOUT_PCT": "24"},
    "carve35": {"LSQR_B200_SMEM_CARVEOUT_PCT": "35"},
    "carve44": {"LSQR_B200_SMEM_CARVEOUT_PCT": "44"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="default,perblock,window")
    ap.add_argument("--workloads", default="C2:1,C3:5,C5:20,C4:10")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--solve", type=int, default=1)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    for wl in args.workloads.split(","):
        name, scale = wl.split(":")
        cfg = synth.scaled(name, float(scale))
        m, n = cfg["m"], cfg["n"]
        irow, icol, a = synth_device.coo_block(cfg["kind"], cfg["seed"], m, n, cfg["k"], 0, m, dev)
        nnz = a.numel()
        xt = synth_device.x_true(cfg["seed"], n, dev)
        yt = synth_device.noise(cfg["seed"], 0, m, dev, scale=1.0)
        ref = {}
        for mode_name in args.modes.split(","):
            for env in MODES.values():
                for k in env:
                    os.environ.pop(k, None)
            os.environ.update(MODES[mode_name])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s = lsqr_b200.LsqrSolverEz().initialize(m, n, a, irow, icol, stream=stream,
                                                    atol=1e-10, btol=1e-10, conlim=1e8, itnlim=100000)
            torch.cuda.synchronize()
            t_init = time.perf_counter() - t0
            pa, pat = s.plan(False), s.plan(True)
            out = {"workload": f"{name}/{scale}", "m": m, "n": n, "nnz": nnz, "mode": mode_name, "init_s": round(t_init, 3),
                   "blocks": [pa["nblocks"], pat["nblocks"]], "window": [pa["window_doubles"], pat["window_doubles"]],
                   "windowed": [round(pa["windowed_fraction"], 3), round(pat["windowed_fraction"], 3)],
                   "ctas_per_sm": [pa["ctas_per_sm"], pat["ctas_per_sm"]], "span_max": [pa["span_max"], pat["span_max"]],
                   "epl": [pa["entries_per_lane"], pat["entries_per_lane"]],
                   "lines": [round(pa["lines_per_gather"], 1), round(pat["lines_per_gather"], 1)]}
            for mode in (1, 2):
                x = xt.clone()
                y = yt.clone()
                s.aprod(mode, m, n, x, y)
                res = (y if mode == 1 else x).clone()
                key = f"mode{mode}"
                if key in ref:
                    out[f"{key}_maxrel_vs_first"] = float((res - ref[key]).abs().max() / ref[key].abs().max())
                else:
                    ref[key] = res
                for _ in range(3):
                    s.aprod_device(mode, m, n, x, y, stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    s.aprod_device(mode, m, n, x, y, stream)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / args.reps
                rows, cols = (m, n) if mode == 1 else (n, m)
                byts = 12 * nnz + 4 * (rows + 1) + 8 * cols + 16 * rows
                out[f"{key}_us"] = round(us, 2)
                out[f"{key}_GBps"] = round(byts / us / 1e3, 1)
                out[f"{key}_frac"] = round(byts / us / 1e3 / peak, 4)
            # the real loop alternates A and A': time the two kernels back to back, each between its own events
            x, y = xt.clone(), yt.clone()
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.reps)]
            for _ in range(2):
                s.aprod_device(1, m, n, x, y, stream); s.aprod_device(2, m, n, x, y, stream)
            for e in ev:
                e[0].record(); s.aprod_device(1, m, n, x, y, stream)
                e[1].record(); s.aprod_device(2, m, n, x, y, stream)
                e[2].record()
            torch.cuda.synchronize()
            a1 = sorted(e[0].elapsed_time(e[1]) * 1e3 for e in ev)[len(ev) // 2]
            a2 = sorted(e[1].elapsed_time(e[2]) * 1e3 for e in ev)[len(ev) // 2]
            out["alt_mode1_us"], out["alt_mode2_us"] = round(a1, 2), round(a2, 2)
            out["alt_frac"] = round((24 * nnz + 28 * m + 28 * n) / ((a1 + a2) * 1e-6) / 1e9 / peak, 4)
            if args.solve:
                b = yt.clone()
                s.aprod(1, m, n, xt.clone(), b)
                xo = torch.empty(n, dtype=torch.float64, device=dev)
                s.solve(b, cfg["damp"], x=xo)
                r = s.solve(b, cfg["damp"], x=xo)
                kt = s.kernel_times()
                biter = synth.b_iter_bytes(nnz, m, n)
                out.update(itn=r.itn, istop=r.istop, loop_ms=round(kt["loop_ms"], 3),
                           us_per_iter=round(1e3 * kt["loop_ms"] / max(r.itn, 1), 2),
                           loop_frac=round(biter * r.itn / (kt["loop_ms"] * 1e-3) / 1e9 / peak, 4))
                # true cost of one iteration: slope of the loop time between two fixed iteration counts
                for graph in (1, 0):
                    ms = {}
                    for lim in (32, 96):
                        s.set_tolerances(atol=0.0, btol=0.0, conlim=0.0, itnlim=lim, use_graph=bool(graph))
                        s.solve(b, cfg["damp"], x=xo)
                        rr = s.solve(b, cfg["damp"], x=xo)
                        ms[lim] = s.kernel_times()["loop_ms"] if rr.itn == lim else float("nan")
                    us = 1e3 * (ms[96] - ms[32]) / 64
                    out["slope_us_graph" if graph else "slope_us_nograph"] = round(us, 2)
                    if graph:
                        out["slope_frac"] = round(biter / (us * 1e-6) / 1e9 / peak, 4)
                s.set_tolerances(atol=1e-10, btol=1e-10, conlim=1e8, itnlim=100000, use_graph=True)
                s.solve(b, cfg["damp"], x=xo)
                key = "x"
                xs = xo.clone()
                if key in ref:
                    out["x_rel_vs_first"] = float((xs - ref[key]).norm() / ref[key].norm())
                else:
                    ref[key] = xs
            print(json.dumps(out), flush=True)
            del s
        del irow, icol, a
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

#!/bin/bash
# Round 2, run 3 (1 GPU): cooperative guarded launches + exact carve-out steps + lane-consecutive gathers.
mkdir -p gpurun_out
echo "== quick guard check (a blocked, windowed problem: the case that hung run 2)"
timeout 120 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "kernel_modes and C3 and smallwindow" 2>&1 | tail -3
echo "== spmv A/B"; timeout 600 python scripts/spmv_bench.py --modes default,blockedgather,nostriped,widewindow,perblock --workloads C3:1,C5:4,C2:1 --reps 10 > gpurun_out/spmv_bench.jsonl 2> gpurun_out/spmv_bench.err; echo "spmv rc=$?"
python - <<'P'
import json
for l in open("gpurun_out/spmv_bench.jsonl"):
    d = json.loads(l)
    print({k: d.get(k) for k in ("workload", "mode", "blocks", "window", "striped", "lines", "ctas_per_sm", "mode1_us", "mode1_frac", "mode2_us", "mode2_frac", "alt_frac", "us_per_iter", "loop_frac", "slope_us_graph", "itn", "x_rel_vs_first")})
P
tail -5 gpurun_out/spmv_bench.err
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"; LSQR_B200_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "cold", {k: d["e2e_cold"][k] for k in ("initialize_s", "first_solve_s")}, "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "loop_frac")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("clocks", d["clocks"]); print("check", d["check"])
    for s in d.get("secondary", []): print("secondary", {k: s[k] for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline", "itn_per_step")}, s["roofline"]["per_kernel"], s["plan"])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
P
grep -v "^\[lsqr_b200 trace\]" gpurun_out/bench_default.err | tail -12 | cut -c1-300

#!/bin/bash
# Round 2, run 1 (1 GPU): smoke, the whole GPU suite (no -x: every failure listed), gather microbench, kernel A/B, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; LSQR_B200_VERBOSE=1 timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== gather"; timeout 200 ./build/gather_bench > gpurun_out/gather_bench.txt 2>&1; echo "gather rc=$?"; cat gpurun_out/gather_bench.txt | cut -c1-200
echo "== spmv A/B"; timeout 600 python scripts/spmv_bench.py --modes default,perblock,nowindow,noguard --workloads C2:1,C3:1,C4:2,C5:4 --reps 10 > gpurun_out/spmv_bench.jsonl 2> gpurun_out/spmv_bench.err; echo "spmv rc=$?"
python - <<'P'
import json
for l in open("gpurun_out/spmv_bench.jsonl"):
    d = json.loads(l)
    print({k: d.get(k) for k in ("workload", "mode", "blocks", "window", "windowed", "ctas_per_sm", "mode1_us", "mode1_frac", "mode2_us", "mode2_frac", "alt_frac", "us_per_iter", "loop_frac", "slope_us_graph", "slope_frac", "itn", "x_rel_vs_first")})
P
tail -5 gpurun_out/spmv_bench.err
echo "== bench"; LSQR_B200_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "cold", d["e2e_cold"], "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "loop_frac")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("clocks", d["clocks"]); print("check", d["check"])
    for s in d.get("secondary", []): print("secondary", s)
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
P
grep -v "^\[lsqr_b200 trace\]" gpurun_out/bench_default.err | tail -25 | cut -c1-300
ls -la gpurun_out | tail -12

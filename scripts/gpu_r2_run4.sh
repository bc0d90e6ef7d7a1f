#!/bin/bash
# Round 2, run 4 (1 GPU): residency probe instead of cooperative launches, inline rare ssq path, CTA-level guard polling.
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== spmv A/B"; timeout 600 python scripts/spmv_bench.py --modes default,epl4,epl8,epl8nowindow,epl8win640,perblock --workloads C3:1,C5:4,C2:1,C4:2 --reps 10 > gpurun_out/spmv_bench.jsonl 2> gpurun_out/spmv_bench.err; echo "spmv rc=$?"
python - <<'P'
import json
for l in open("gpurun_out/spmv_bench.jsonl"):
    d = json.loads(l)
    print({k: d.get(k) for k in ("workload", "mode", "blocks", "window", "epl", "ctas_per_sm", "mode1_us", "mode1_frac", "mode2_us", "mode2_frac", "alt_frac", "us_per_iter", "loop_frac", "slope_us_graph", "itn", "x_rel_vs_first")})
P
tail -5 gpurun_out/spmv_bench.err
echo "== bench"; LSQR_B200_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "cold", {k: d["e2e_cold"][k] for k in ("initialize_s", "first_solve_s")}, "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "loop_frac")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("clocks", d["clocks"]); print("check", d["check"]["oracle"])
    for s in d.get("secondary", []): print("secondary", {k: s[k] for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline", "itn_per_step")}, s["roofline"]["per_kernel"], s["plan"])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
P
grep -v "^\[lsqr_b200 trace\]" gpurun_out/bench_default.err | grep -v "^\[lsqr_b200\] " | tail -8 | cut -c1-300
echo "== ncu full: C3 kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 30 -c 4 -f -o gpurun_out/prof_c3 \
   python bench.py --workload C3 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log | cut -c1-200
echo "== ncu full: C5/4 kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 30 -c 4 -f -o gpurun_out/prof_c5q \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu_c5q.log 2>&1; tail -2 gpurun_out/ncu_c5q.log | cut -c1-200
echo "== trace C5"; LSQR_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c5.txt > /dev/null; grep "trace\]" gpurun_out/trace_c5.txt | tail -30
ls -la gpurun_out | tail -8

#!/bin/bash
# Round-1 second verification pass (1 GPU): new tests, default bench (C5 + secondary C2), reference arm, ncu evidence for C5.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "balanced or two_gpu or ez_kats or blocked" 2>&1 | tail -4
LSQR_B200_VERBOSE=1 timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"; python - <<'P'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print({k: d[k] for k in ("value", "ms_per_step", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline", "gpu_launches")})
print("e2e", d["e2e"]["value"], "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "loop_frac")})
print("per_kernel", d["roofline"]["per_kernel"]); print("clocks", d["clocks"]); print("secondary", d.get("secondary")); print("cpu", d.get("cpu_baseline"))
P
tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
# every launch of two iterations of the real C5 loop with its DRAM bytes (single pass: no replay)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 60 --csv \
   --log-file gpurun_out/c5_launches_dram.csv python bench.py --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph > gpurun_out/ncu_c5_launch.log 2>&1
tail -2 gpurun_out/ncu_c5_launch.log
# full sections of the SpMV kernel on C5 at 1/4 scale (kernel replay has to save / restore the written buffers)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_warp -s 200 -c 6 -f -o gpurun_out/prof_c5q \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph > gpurun_out/ncu_c5q.log 2>&1
tail -2 gpurun_out/ncu_c5q.log
ls -la gpurun_out | tail -12

#!/bin/bash
# Full verification pass on one B200: GPU parity tests, smoke, bench, kernel A/B bench, device timeline, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
tail -45 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
LSQR_B200_VERBOSE=1 timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
LSQR_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/trace_default.err
grep "trace" gpurun_out/trace_default.err | tail -50 > gpurun_out/trace_default_timeline.txt
timeout 900 python scripts/spmv_bench.py --variants 3 --workloads C2:1,C3:2,C5:4,C4:4 > gpurun_out/spmv_bench.jsonl 2> gpurun_out/spmv_bench.err
cat gpurun_out/spmv_bench.jsonl; tail -5 gpurun_out/spmv_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_warp -s 40 -c 4 -f -o gpurun_out/prof_default python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

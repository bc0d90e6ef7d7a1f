#!/bin/bash
# One GPU session: parity tests, smoke, kernel A/B bench, bench, ncu.  Logs land in gpurun_out/.
#   scripts/gpu_round.sh [ncu] [quick]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 900 python scripts/spmv_bench.py ${SPMV_BENCH_ARGS} > gpurun_out/spmv_bench.jsonl 2> gpurun_out/spmv_bench.err
cat gpurun_out/spmv_bench.jsonl; tail -5 gpurun_out/spmv_bench.err
LSQR_B200_VERBOSE=1 timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
cat gpurun_out/bench_c2.json; tail -8 gpurun_out/bench_c2.err
if [[ " $* " == *" ncu "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 40 -c 4 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi

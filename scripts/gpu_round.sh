#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
LSQR_B200_VERBOSE=1 timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 40 -c 4 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
fi
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_c2.json; tail -8 gpurun_out/bench_c2.err

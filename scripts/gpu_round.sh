#!/bin/bash
# One GPU session: parity tests, smoke, bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt; lscpu | head -20 >> gpurun_out/host_mem.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err

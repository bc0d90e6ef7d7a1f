#!/bin/bash
# Round 2, run 8 (1 GPU): the kernel flavour without the window path (tests + same-box A/B against the round-1 tree),
# ncu --set full of the C5/4 products of both trees (what differs?), DRAM bytes per launch of one full-size C5 iteration.
mkdir -p gpurun_out
show() {
python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ("workload", "mode", "variant", "mode1_us", "mode2_us", "alt_mode1_us", "alt_mode2_us", "us_per_iter", "loop_frac")})
P
}
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or log_lines or stream or hook or C5" > gpurun_out/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu_subset.log | cut -c1-300
for rep in 1 2; do
  echo "== A/B rep $rep: r01"
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4 --reps 10) > gpurun_out/ab8_r01_$rep.jsonl 2> gpurun_out/ab8_r01_$rep.err; echo "rc=$?"; show gpurun_out/ab8_r01_$rep.jsonl
  echo "== A/B rep $rep: r02 (no-window flavour)"
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4 --reps 10 > gpurun_out/ab8_r02_$rep.jsonl 2> gpurun_out/ab8_r02_$rep.err; echo "rc=$?"; show gpurun_out/ab8_r02_$rep.jsonl
done
echo "== ncu full: C5/4, r01 tree"
(cd build/r01tree && timeout 600 ncu --set full --clock-control none -k regex:spmv_warp_kernel -s 30 -c 2 -f -o ../../gpurun_out/prof_c5q_r01 \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph) > gpurun_out/ncu_c5q_r01.log 2>&1; tail -2 gpurun_out/ncu_c5q_r01.log | cut -c1-200
echo "== ncu full: C5/4, r02"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 30 -c 2 -f -o gpurun_out/prof_c5q_r02 \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu_c5q_r02.log 2>&1; tail -2 gpurun_out/ncu_c5q_r02.log | cut -c1-200
echo "== ncu launch list with DRAM bytes: one full-size C5 iteration"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"spmv_kernel|xw_update|nrm2|init_w" -c 14 --csv --log-file gpurun_out/c5_launches_dram.csv \
   python bench.py --steps 1 --warmup 1 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu_c5_launches.log 2>&1; tail -2 gpurun_out/ncu_c5_launches.log | cut -c1-200
grep -c spmv_kernel gpurun_out/c5_launches_dram.csv
ls -la gpurun_out | tail -8

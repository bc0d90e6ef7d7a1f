#!/bin/bash
# Round 2, run 11 (1 GPU): 128-byte aligned shared segment -- same-box A/B against the round-1 tree, ncu --set full of both.
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
for rep in 1 2 3; do
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C2:1 --reps 10) > gpurun_out/ab11_r01_$rep.jsonl 2>/dev/null; show gpurun_out/ab11_r01_$rep.jsonl
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C2:1 --reps 10 > gpurun_out/ab11_r02_$rep.jsonl 2>/dev/null; show gpurun_out/ab11_r02_$rep.jsonl
done
echo "== ncu full: C5/4, r01 tree"
(cd build/r01tree && timeout 600 ncu --set full --clock-control none -k regex:spmv_warp_kernel -s 30 -c 2 -f -o ../../gpurun_out/prof11_c5q_r01 \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph) > gpurun_out/ncu11_c5q_r01.log 2>&1; tail -1 gpurun_out/ncu11_c5q_r01.log | cut -c1-200
echo "== ncu full: C5/4, r02"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 30 -c 2 -f -o gpurun_out/prof11_c5q_r02 \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu11_c5q_r02.log 2>&1; tail -1 gpurun_out/ncu11_c5q_r02.log | cut -c1-200

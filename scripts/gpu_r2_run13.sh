#!/bin/bash
# Round 2, run 13 (1 GPU): first use of the gathered values behind the head mask (as the round-1 kernel had it), scan
# steps skipped when no segment reaches that far -- tests, same-box A/B (round-1 tree, this tree, this tree with all scan steps).
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
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or stream or aprod or csr or tile" > gpurun_out/pytest_gpu_subset13.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset13.log | cut -c1-300
for rep in 1 2; do
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C3:1,C2:1,C4:2 --reps 10) > gpurun_out/ab13_r01_$rep.jsonl 2>/dev/null; show gpurun_out/ab13_r01_$rep.jsonl
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab13_r02_$rep.jsonl 2>/dev/null; show gpurun_out/ab13_r02_$rep.jsonl
  LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.fullscan.so timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab13_r02fullscan_$rep.jsonl 2>/dev/null; show gpurun_out/ab13_r02fullscan_$rep.jsonl
done

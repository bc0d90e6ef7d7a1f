#!/bin/bash
# Round 2, run 12 (1 GPU): row-sum buffer addressed in the shared state space with a pinned 32-bit address -- tests and
# same-box A/B against the round-1 tree.
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
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or stream or aprod or csr or tile" > gpurun_out/pytest_gpu_subset12.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset12.log | cut -c1-300
for rep in 1 2 3; do
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C3:1,C2:1 --reps 10) > gpurun_out/ab12_r01_$rep.jsonl 2>/dev/null; show gpurun_out/ab12_r01_$rep.jsonl
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1 --reps 10 > gpurun_out/ab12_r02_$rep.jsonl 2>/dev/null; show gpurun_out/ab12_r02_$rep.jsonl
done

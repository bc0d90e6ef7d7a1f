#!/bin/bash
# Round 2, run 10 (1 GPU): which shared-memory carve-out does the requested percentage select?  29 % of 228 KB is 66 KB:
# if the driver rounds UP to the next configuration the kernels run with 100 KB of shared memory and 156 KB of L1
# instead of 64 / 192.  One process per setting; the round-1 tree (no carve-out attribute) beside them.
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
for rep in 1 2; do
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4 --reps 10) > gpurun_out/ab10_r01_$rep.jsonl 2>/dev/null; show gpurun_out/ab10_r01_$rep.jsonl
  for mode in default carve28 carve24 carve35 carve44; do
    timeout 300 python scripts/spmv_bench.py --modes $mode --workloads C5:4 --reps 10 > gpurun_out/ab10_${mode}_$rep.jsonl 2>/dev/null; show gpurun_out/ab10_${mode}_$rep.jsonl
  done
done

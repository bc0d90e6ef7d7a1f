#!/bin/bash
# Round 2, run 16 (1 GPU): tile size / row weight of the work plan after the chunk-loop changes (C5/4, C2, C3).
mkdir -p gpurun_out
show() {
python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ("workload", "mode", "mode1_us", "mode2_us", "alt_mode1_us", "alt_mode2_us", "us_per_iter", "loop_frac")})
P
}
for rep in 1 2; do
  timeout 600 python scripts/spmv_bench.py --modes default,tile4k,tile16k,tile32k,roww2,roww8 --workloads C5:4,C2:1,C3:1 --reps 10 > gpurun_out/ab16_$rep.jsonl 2>/dev/null; show gpurun_out/ab16_$rep.jsonl
done

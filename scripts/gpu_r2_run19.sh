#!/bin/bash
# Round 2, run 19 (1 GPU): iterations per CUDA-graph launch on the small problem (C2; default 3): gap between graphs
# (~15 us) against empty iterations enqueued past the stop.
mkdir -p gpurun_out
for rep in 1 2 3; do
  timeout 300 python scripts/spmv_bench.py --modes default,batch1,batch2,batch5,batch8,batch12 --workloads C2:1 --reps 5 > gpurun_out/ab19_$rep.jsonl 2>/dev/null
  python - gpurun_out/ab19_$rep.jsonl <<'P'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ("workload", "mode", "us_per_iter", "loop_ms", "itn", "slope_us_graph")})
P
done

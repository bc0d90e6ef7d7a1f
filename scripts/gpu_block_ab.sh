#!/bin/bash
# C5 on one GPU: L2 budget of the gathered vector (column blocks of A / row blocks of A')
mkdir -p gpurun_out
for cfg in ${CFGS:-96:48 96:64 48:64 96:80}; do
  IFS=: read vb ub <<< "$cfg"
  LSQR_B200_VBLOCK_MB=$vb LSQR_B200_UBLOCK_MB=$ub timeout 600 python bench.py --steps 2 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/blk_$cfg.json 2> gpurun_out/blk_$cfg.err
  python - $cfg <<'P'
import json, sys
d = json.load(open(f"gpurun_out/blk_{sys.argv[1]}.json"))
print("V:U", sys.argv[1], {k: round(d[k], 4) for k in ("value", "iters_per_s", "ms_per_iteration", "frac_of_hbm_roofline")},
      {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["clocks"]["sm_mhz"], d["roofline"]["launch_note"][:40])
P
done

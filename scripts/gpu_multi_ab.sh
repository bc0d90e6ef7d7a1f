#!/bin/bash
# N-GPU A/B of the pipelined all-reduce: chunks x stream priority x SM reserve (env triples "chunks:prio:reserve")
N=${1:-8}
mkdir -p gpurun_out
for cfg in ${CFGS:-1:1:0 4:0:0 4:1:0 8:1:0}; do
  IFS=: read chunks prio res <<< "$cfg"
  LSQR_B200_COMM_CHUNKS=$chunks LSQR_B200_COMM_PRIORITY=$prio LSQR_B200_COMM_RESERVE_SMS=$res timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
     bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 > gpurun_out/ab_n${N}_$cfg.json 2> gpurun_out/ab_n${N}_$cfg.err
  python - $N $cfg <<'P'
import json, sys
d = json.load(open(f"gpurun_out/ab_n{sys.argv[1]}_{sys.argv[2]}.json"))
print(sys.argv[2], {k: round(d[k], 4) for k in ("value", "iters_per_s", "ms_per_iteration", "frac_of_hbm_roofline")},
      "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["clocks"]["sm_mhz"])
P
done

#!/bin/bash
# Full-size bench lines of the other BASELINE workload families (C3 banded damped, C4 power law) on N GPUs
N=${1:-1}
mkdir -p gpurun_out
for wl in ${WLS:-C3 C4}; do
  if [ "$N" == "1" ]; then
    timeout 600 python bench.py --workload $wl --steps ${STEPS:-5} --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench_${wl}_n1.json 2> gpurun_out/bench_${wl}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  fi
  python - $wl $N <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_n{sys.argv[2]}.json"))
print(sys.argv[1], "N=" + sys.argv[2], {k: round(d[k], 4) for k in ("value", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline")},
      "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["clocks"]["sm_mhz"], d["check"]["ok"], d["config"]["workload"][:60])
P
done

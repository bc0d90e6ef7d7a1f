#!/bin/bash
# Round 2, 8-GPU run: C5 through the NVLink peer-memory exchange and through the NCCL all-reduce path (A/B); each line carries the
# oracle parity check of the same path at reduced scale.
N=${1:-8}
mkdir -p gpurun_out
for peer in 1 0; do
  echo "== bench N=$N peer=$peer"
  LSQR_B200_PEER_EXCHANGE=$peer LSQR_B200_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2966$peer bench.py --gpus $N --steps 10 --warmup 3 --secondary none --no-cpu-baseline \
      > gpurun_out/bench_n${N}_peer$peer.json 2> gpurun_out/bench_n${N}_peer$peer.err; echo "rc=$?"
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_n${N}_peer$peer.json"))
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration", "collective_ms")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("check", d["check"]["oracle"]); print("e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["total_s"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
P
  grep -v "trace\]" gpurun_out/bench_n${N}_peer$peer.err | grep -i "error\|assert\|Traceback\|exchange over\|timed out" | sort | uniq -c | head -12
done

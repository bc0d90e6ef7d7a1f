#!/bin/bash
# Multi-GPU pass: N = $1 ranks.  Parity worker (2 ranks) when N == 2, then the bench at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" == "2" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 tests/mgpu_worker.py 2>&1 | grep -E "MGPU_OK|Error|error|assert" | head -20
fi
for chunks in ${CHUNKS:-4}; do
LSQR_B200_COMM_CHUNKS=$chunks LSQR_B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 \
   bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n${N}_c${chunks}.json 2> gpurun_out/bench_n${N}_c${chunks}.err
echo "bench N=$N chunks=$chunks rc=$?"
python - $N $chunks <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_n{sys.argv[1]}_c{sys.argv[2]}.json"))
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "iters_per_s", "itn_per_step", "ms_per_iteration", "frac_of_hbm_roofline", "gpu_launches")})
print("e2e", d["e2e"]["value"], "roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms", "loop_frac")})
print("per_kernel", {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}); print("clocks", d["clocks"])
P
grep "lsqr_b200\]" gpurun_out/bench_n${N}_c${chunks}.err | head -2
done

#!/bin/bash
# Round 2, multi-GPU run (N GPUs, default 2): the GPU suite (its two-GPU test runs the parity worker: peer-memory exchange and NCCL
# all-reduce paths against the oracle), then the C5 bench through both exchange paths.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
echo "== worker"; LSQR_B200_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29641 tests/mgpu_worker.py > gpurun_out/mgpu_worker_n$N.log 2>&1; echo "worker rc=$?"
grep -c MGPU_OK gpurun_out/mgpu_worker_n$N.log; grep "MGPU_OK\|Error\|error\|assert\|exchange over" gpurun_out/mgpu_worker_n$N.log | sort | uniq -c | head -30
tail -5 gpurun_out/mgpu_worker_n$N.log | cut -c1-300
if [ "$2" = "pytest" ]; then
  echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"
  tail -8 gpurun_out/pytest_gpu_n$N.log | cut -c1-300
fi
for peer in 1 0; do
  echo "== bench N=$N peer=$peer"
  LSQR_B200_PEER_EXCHANGE=$peer LSQR_B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2965$peer bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/bench_n${N}_peer$peer.json 2> gpurun_out/bench_n${N}_peer$peer.err; echo "rc=$?"
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_n${N}_peer$peer.json"))
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration", "collective_ms")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("check", d["check"]["oracle"]); print("e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["total_s"])
except Exception as e:
    print("no bench line:", e)
P
  grep -v "trace\]" gpurun_out/bench_n${N}_peer$peer.err | grep -i "error\|assert\|Traceback\|exchange over" | sort | uniq -c | head -10
done

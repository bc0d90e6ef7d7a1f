#!/bin/bash
# Round 2, final multi-GPU pass: N = 2 runs the whole GPU suite (its two-GPU test drives tests/mgpu_worker.py: 6 cases,
# both exchange paths, against the serial oracle) and the C5 bench through the peer-memory exchange; N = 4 the bench only.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"
  tail -6 gpurun_out/pytest_gpu_n$N.log | cut -c1-300
fi
echo "== bench N=$N (peer-memory exchange)"
LSQR_B200_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err; echo "rc=$?"
python - <<P
import json
try:
    d = json.load(open("gpurun_out/bench_final_n$N.json"))
    print({k: d[k] for k in ("value", "n_gpus", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration", "collective_ms")})
    print("per_kernel", d["roofline"]["per_kernel"]); print("check", d["check"]["oracle"]); print("e2e", d["e2e"]["value"], "cold", d["e2e_cold"]["total_s"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
P
grep -v "trace\]" gpurun_out/bench_final_n$N.err | grep -i "error\|assert\|Traceback\|exchange over\|timed out" | sort | uniq -c | head -10

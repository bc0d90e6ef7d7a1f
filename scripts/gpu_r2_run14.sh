#!/bin/bash
# Round 2, run 14 (1 GPU): kernel flavours chosen per plan (local / gather-bound) -- tests, same-box A/B against the
# round-1 tree on all four families, full-size C5 solve of both trees.
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
benchline() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "per_kernel", d["roofline"]["per_kernel"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line:", e)
P
}
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or stream or aprod or csr or tile" > gpurun_out/pytest_gpu_subset14.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset14.log | cut -c1-300
for rep in 1 2; do
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C3:1,C2:1,C4:2 --reps 10) > gpurun_out/ab14_r01_$rep.jsonl 2>/dev/null; show gpurun_out/ab14_r01_$rep.jsonl
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab14_r02_$rep.jsonl 2>/dev/null; show gpurun_out/ab14_r02_$rep.jsonl
done
timeout 300 python scripts/spmv_bench.py --modes local,gather --workloads C3:1,C4:2 --reps 10 > gpurun_out/ab14_r02_flavours.jsonl 2>/dev/null; show gpurun_out/ab14_r02_flavours.jsonl
echo "== full C5, r01 tree"
(cd build/r01tree && timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline) > gpurun_out/bench14_c5_r01tree.json 2> gpurun_out/bench14_c5_r01tree.err; echo "rc=$?"; benchline gpurun_out/bench14_c5_r01tree.json
echo "== full C5, r02"
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench14_c5_r02.json 2> gpurun_out/bench14_c5_r02.err; echo "rc=$?"; benchline gpurun_out/bench14_c5_r02.json

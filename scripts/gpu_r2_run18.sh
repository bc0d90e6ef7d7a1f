#!/bin/bash
# Round 2, run 18 (1 GPU): chunks without a row end skip the shared-memory round trip (liblsqr_b200.skipsts.so) against the
# current default: tests under the new build, products and solves of all families, full-size C5.
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
benchline() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "per_kernel", d["roofline"]["per_kernel"], "clocks", d["clocks"], "oracle ok", d["check"]["oracle"]["ok"])
except Exception as e:
    print("no bench line:", e)
P
}
H=$PWD/lsqr_b200/lib/liblsqr_b200.skipsts.so
echo "== pytest subset (skip build)"; LSQR_B200_LIB=$H timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or stream or aprod or csr or tile or empty or ragged or single or duplicate or long_rows or flavour" > gpurun_out/pytest_gpu_subset18.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset18.log | cut -c1-300
for rep in 1 2; do
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab18_default_$rep.jsonl 2>/dev/null; show gpurun_out/ab18_default_$rep.jsonl
  LSQR_B200_LIB=$H timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab18_skipsts_$rep.jsonl 2>/dev/null; show gpurun_out/ab18_skipsts_$rep.jsonl
done
echo "== full C5, default"
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench18_c5_default.json 2>/dev/null; echo "rc=$?"; benchline gpurun_out/bench18_c5_default.json
echo "== full C5, skip build"
LSQR_B200_LIB=$H timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench18_c5_skipsts.json 2>/dev/null; echo "rc=$?"; benchline gpurun_out/bench18_c5_skipsts.json

#!/bin/bash
# Round 2, run 15 (1 GPU): row-end sums by shuffle when at most two rows end in a chunk (liblsqr_b200.hyb.so) against the
# current default, tests under the new build.
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
H=$PWD/lsqr_b200/lib/liblsqr_b200.hyb.so
echo "== pytest subset (hybrid build)"; LSQR_B200_LIB=$H timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or stream or aprod or csr or tile or empty or ragged or single or duplicate" > gpurun_out/pytest_gpu_subset15.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset15.log | cut -c1-300
for rep in 1 2; do
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab15_default_$rep.jsonl 2>/dev/null; show gpurun_out/ab15_default_$rep.jsonl
  LSQR_B200_LIB=$H timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1,C2:1,C4:2 --reps 10 > gpurun_out/ab15_hyb_$rep.jsonl 2>/dev/null; show gpurun_out/ab15_hyb_$rep.jsonl
done

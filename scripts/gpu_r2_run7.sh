#!/bin/bash
# Round 2, run 7 (1 GPU): full GPU suite on the final kernel configuration, then a same-box A/B of the round-1 tree
# (build/r01tree, commit 055d1ab) against this tree and against a build without the scaling test of the sums of
# squares, to separate box-to-box variance from a real change in the kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt; cat gpurun_out/gpu.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
show() {
python - "$1" <<'P'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print({k: d.get(k) for k in ("workload", "mode", "variant", "mode1_us", "mode2_us", "alt_mode1_us", "alt_mode2_us", "us_per_iter", "loop_frac")})
P
}
for rep in 1 2; do
  echo "== A/B rep $rep: r01"
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C2:1 --reps 10) > gpurun_out/ab_r01_$rep.jsonl 2> gpurun_out/ab_r01_$rep.err; echo "rc=$?"; show gpurun_out/ab_r01_$rep.jsonl
  echo "== A/B rep $rep: r02"
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C2:1 --reps 10 > gpurun_out/ab_r02_$rep.jsonl 2> gpurun_out/ab_r02_$rep.err; echo "rc=$?"; show gpurun_out/ab_r02_$rep.jsonl
  echo "== A/B rep $rep: r02 without the scaling test"
  LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.plainssq.so timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C2:1 --reps 10 > gpurun_out/ab_r02plain_$rep.jsonl 2> gpurun_out/ab_r02plain_$rep.err; echo "rc=$?"; show gpurun_out/ab_r02plain_$rep.jsonl
done
benchline() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    print("e2e", d["e2e"]["value"], "per_kernel", d["roofline"]["per_kernel"], "clocks", d["clocks"])
    print("check", d.get("check")); print("cold", d.get("e2e_cold")); print("cpu", d.get("cpu_baseline"))
    for s in d.get("secondary") or []:
        if isinstance(s, dict): print("secondary", {k: s.get(k) for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline", "itn_per_step")}, s["roofline"].get("per_kernel"))
except Exception as e:
    print("no bench line:", e)
P
}
echo "== full C5, r01 tree"
(cd build/r01tree && timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline) > gpurun_out/bench_c5_r01tree.json 2> gpurun_out/bench_c5_r01tree.err; echo "rc=$?"; benchline gpurun_out/bench_c5_r01tree.json
echo "== full C5, r02"
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench_c5_r02.json 2> gpurun_out/bench_c5_r02.err; echo "rc=$?"; benchline gpurun_out/bench_c5_r02.json
echo "== full C5, r02 without the scaling test"
LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.plainssq.so timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench_c5_r02plain.json 2> gpurun_out/bench_c5_r02plain.err; echo "rc=$?"; benchline gpurun_out/bench_c5_r02plain.json
echo "== full C5, r02 per-block launches"
LSQR_B200_SINGLE_LAUNCH=0 timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check > gpurun_out/bench_c5_r02perblock.json 2> gpurun_out/bench_c5_r02perblock.err; echo "rc=$?"; benchline gpurun_out/bench_c5_r02perblock.json
echo "== C2 scalar step A/B (warp-parallel vs one thread)"
for rep in 1 2; do
  timeout 300 python scripts/spmv_bench.py --modes default,pdl --workloads C2:1 --reps 20 > gpurun_out/ab_c2_warpstep_$rep.jsonl 2>/dev/null; show gpurun_out/ab_c2_warpstep_$rep.jsonl
  LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.serialstep.so timeout 300 python scripts/spmv_bench.py --modes default --workloads C2:1 --reps 20 > gpurun_out/ab_c2_serialstep_$rep.jsonl 2>/dev/null; show gpurun_out/ab_c2_serialstep_$rep.jsonl
done
LSQR_B200_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c2_warpstep.txt > /dev/null; grep "trace\]" gpurun_out/trace_c2_warpstep.txt | sed -n 2,14p
LSQR_B200_PDL=1 LSQR_B200_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c2_pdl.txt > /dev/null; grep "trace\]" gpurun_out/trace_c2_pdl.txt | sed -n 2,14p
LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.serialstep.so LSQR_B200_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c2_serialstep.txt > /dev/null; grep "trace\]" gpurun_out/trace_c2_serialstep.txt | sed -n 2,14p
echo "== C2 bench with PDL"
LSQR_B200_PDL=1 timeout 600 python bench.py --workload C2 --secondary none --no-cpu-baseline > gpurun_out/bench_c2_pdl.json 2> gpurun_out/bench_c2_pdl.err; echo "rc=$?"; benchline gpurun_out/bench_c2_pdl.json
timeout 600 python bench.py --workload C2 --secondary none --no-cpu-baseline > gpurun_out/bench_c2_default.json 2> gpurun_out/bench_c2_default.err; echo "rc=$?"; benchline gpurun_out/bench_c2_default.json
echo "== hook bench C2"
timeout 600 python bench.py --workload C2 --via-hook --secondary none --no-cpu-baseline > gpurun_out/bench_c2_hook.json 2> gpurun_out/bench_c2_hook.err; echo "rc=$?"; benchline gpurun_out/bench_c2_hook.json
tail -3 gpurun_out/*.err | cut -c1-300 | tail -40
ls -la gpurun_out | tail -12

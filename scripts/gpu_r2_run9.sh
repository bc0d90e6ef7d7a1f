#!/bin/bash
# Round 2, run 9 (1 GPU): block mode as a compile-time parameter of the chunk loop -- tests, same-box A/B against the
# round-1 tree (C5/4 products, full C5 solve), dynamic instruction count of the A v kernel.
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
    print("check", d.get("check"))
    for s in d.get("secondary") or []:
        if isinstance(s, dict): print("secondary", {k: s.get(k) for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline", "itn_per_step")}, s["roofline"].get("per_kernel"))
except Exception as e:
    print("no bench line:", e)
P
}
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kernel_modes or window or blocked or kat or readme or log_lines or stream or hook or C5 or aprod or csr" > gpurun_out/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu_subset.log | cut -c1-300
for rep in 1 2; do
  echo "== A/B rep $rep: r01"
  (cd build/r01tree && timeout 300 python scripts/spmv_bench.py --variants 3 --workloads C5:4,C3:1 --reps 10) > gpurun_out/ab9_r01_$rep.jsonl 2> gpurun_out/ab9_r01_$rep.err; echo "rc=$?"; show gpurun_out/ab9_r01_$rep.jsonl
  echo "== A/B rep $rep: r02"
  timeout 300 python scripts/spmv_bench.py --modes default --workloads C5:4,C3:1 --reps 10 > gpurun_out/ab9_r02_$rep.jsonl 2> gpurun_out/ab9_r02_$rep.err; echo "rc=$?"; show gpurun_out/ab9_r02_$rep.jsonl
done
echo "== full C5, r01 tree"
(cd build/r01tree && timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline) > gpurun_out/bench9_c5_r01tree.json 2> gpurun_out/bench9_c5_r01tree.err; echo "rc=$?"; benchline gpurun_out/bench9_c5_r01tree.json
echo "== full C5, r02"
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline > gpurun_out/bench9_c5_r02.json 2> gpurun_out/bench9_c5_r02.err; echo "rc=$?"; benchline gpurun_out/bench9_c5_r02.json
echo "== instruction counts, C5/4"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:spmv_kernel -s 30 -c 2 --csv --log-file gpurun_out/inst_c5q_r02.csv \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > /dev/null 2>&1; grep spmv gpurun_out/inst_c5q_r02.csv | cut -d, -f5,13- | head
ls -la gpurun_out | tail -5

#!/bin/bash
# Round 2, final verification pass (1 GPU): smoke, full GPU suite, default bench line, reference arm, hook line, C4 line,
# ncu launch list of the C5 loop, ncu --set full of the C5/4 products, device timelines of C2 and C3.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt; cat gpurun_out/gpu.txt
benchline() {
python - "$1" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("impl", "value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "gpu_launches", "launches_per_iteration")})
    if "roofline" in d: print("e2e", d["e2e"]["value"], "per_kernel", d["roofline"]["per_kernel"], "traffic", d["roofline"].get("traffic"), "clocks", d["clocks"])
    print("check", d.get("check")); print("cold", d.get("e2e_cold")); print("cpu", d.get("cpu_baseline"))
    for s in d.get("secondary") or []:
        if isinstance(s, dict): print("secondary", {k: s.get(k) for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline", "itn_per_step")}, s["roofline"].get("per_kernel"))
except Exception as e:
    print("no bench line:", e)
P
}
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench default"; LSQR_B200_VERBOSE=1 timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; benchline gpurun_out/bench_default.json
grep "flavour\|single_launch" gpurun_out/bench_default.err | sort | uniq -c | cut -c1-260 | head -12
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; benchline gpurun_out/bench_reference.json
echo "== bench C2 via hook"; timeout 600 python bench.py --workload C2 --via-hook --secondary none --no-cpu-baseline > gpurun_out/bench_c2_hook.json 2> gpurun_out/bench_c2_hook.err; echo "rc=$?"; benchline gpurun_out/bench_c2_hook.json
echo "== bench C4"; timeout 900 python bench.py --workload C4 --secondary none --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "rc=$?"; benchline gpurun_out/bench_c4.json
echo "== ncu launch list (C5 loop)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c5_launches.csv \
   python bench.py --steps 2 --warmup 1 --secondary none --no-cpu-baseline --no-oracle-check > gpurun_out/ncu_launches.log 2>&1; grep -c "spmv_kernel" gpurun_out/c5_launches.csv
echo "== ncu full: C5/4 products"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 30 -c 2 -f -o gpurun_out/prof_final_c5q \
   python bench.py --workload C5 --scale 4 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-graph --no-oracle-check > gpurun_out/ncu_final_c5q.log 2>&1; tail -1 gpurun_out/ncu_final_c5q.log | cut -c1-200
echo "== traces"
LSQR_B200_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c2.txt > /dev/null; grep "trace\]" gpurun_out/trace_c2.txt | sed -n 2,9p
LSQR_B200_TRACE=1 timeout 300 python bench.py --workload C3 --steps 1 --warmup 3 --secondary none --no-cpu-baseline --no-oracle-check 2> gpurun_out/trace_c3.txt > /dev/null; grep "trace\]" gpurun_out/trace_c3.txt | sed -n 2,9p; grep "trace\]" gpurun_out/trace_c3.txt | tail -6
ls -la gpurun_out | tail -14

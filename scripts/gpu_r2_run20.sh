#!/bin/bash
# Round 2, run 20 (1 GPU): up to 8 iterations per graph launch -- the tests that look at iteration counts, logs, stopping
# rules and repeated solves, then the C2 and default bench lines.
mkdir -p gpurun_out
echo "== pytest subset"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "kat or readme or log or itnlim or stop or graph or repeated or determin or se_ or wantse or damp or kernel_modes or lstp or cpp" > gpurun_out/pytest_gpu_subset20.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_subset20.log | cut -c1-300
timeout 600 python bench.py --workload C2 --secondary none --no-cpu-baseline > gpurun_out/bench20_c2.json 2>/dev/null; echo "rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench20_c2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration")}, "e2e", d["e2e"]["value"], "oracle ok", d["check"]["oracle"]["ok"])
P

#!/bin/bash
# Round 2, last check of HEAD (1 GPU): whole GPU suite, smoke, C5 bench line without the CPU leg.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_head.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke_head.log | cut -c1-200
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/pytest_gpu_head.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_head.log | cut -c1-300
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_head.json 2>/dev/null; echo "rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_head.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iteration", "itn_per_step", "frac_of_hbm_roofline", "launches_per_iteration")}, "e2e", d["e2e"]["value"], "oracle ok", d["check"]["oracle"]["ok"])
for s in d.get("secondary") or []: print({k: s.get(k) for k in ("workload", "value", "ms_per_iteration", "frac_of_hbm_roofline")})
P

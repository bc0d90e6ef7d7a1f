#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "balanced or ragged or long_rows or scaled_configs" 2>&1 | tail -5
for bal in 0 1; do
  echo "== LSQR_B200_BALANCE=$bal"
  LSQR_B200_VERBOSE=1 LSQR_B200_BALANCE=$bal timeout 600 python scripts/spmv_bench.py --variants 3 --workloads C4:1,C4:8,C2:1 --reps 20 2> gpurun_out/lpt_$bal.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['workload'], 'nnz', d['nnz'], 'm1', d['mode1_us'], d['mode1_frac'], 'm2', d['mode2_us'], d['mode2_frac'], 'iter_us', d['us_per_iter'], 'loop_frac', d['loop_frac'], 'itn', d['itn'], 'init_s', d['init_s'])
" | tee gpurun_out/lpt_$bal.txt
  grep "lsqr_b200\]" gpurun_out/lpt_$bal.err | tail -3
done

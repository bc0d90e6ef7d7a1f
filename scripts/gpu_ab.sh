#!/bin/bash
# A/B of experimental library builds (lsqr_b200/lib/liblsqr_b200.<tag>.so): kernel bench per tag, then parity tests on $PARITY_TAG
mkdir -p gpurun_out
WL=${WL:-C2:1,C3:4,C5:8,C4:8}
for tag in $TAGS; do
  echo "== $tag"
  LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.$tag.so timeout 600 python scripts/spmv_bench.py --variants 3 --workloads $WL --reps 20 2> gpurun_out/ab_$tag.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['workload'], 'm1', d['mode1_us'], d['mode1_frac'], 'm2', d['mode2_us'], d['mode2_frac'], 'iter_us', d['us_per_iter'], 'loop_frac', d['loop_frac'], 'itn', d['itn'])
" | tee gpurun_out/ab_$tag.txt
  tail -2 gpurun_out/ab_$tag.err
done
if [ -n "$PARITY_TAG" ]; then
  LSQR_B200_LIB=$PWD/lsqr_b200/lib/liblsqr_b200.$PARITY_TAG.so timeout 900 python -m pytest tests -m gpu -q -x -k "not other_variants and not full_size" 2>&1 | tail -5
fi

timeout 900 python -m pytest tests -m gpu -q -x -k "blocked or balanced or full_size" 2>&1 | tail -3
for f in 0 1; do
LSQR_B200_FUSE_LAST_BLOCK=$f timeout 600 python bench.py --steps 3 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('fuse_last=$f', {k: round(d[k], 4) for k in ('value', 'iters_per_s', 'ms_per_iteration', 'frac_of_hbm_roofline')}, {k: round(v['ms'], 4) for k, v in d['roofline']['per_kernel'].items()}, d['clocks']['sm_mhz'], d['gpu_launches'])"
done

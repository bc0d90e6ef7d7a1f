mkdir -p gpurun_out
(while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; sleep 2; done) > gpurun_out/c5_mem.txt 2>&1 &
MP=$!
LSQR_B200_VERBOSE=1 timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err
echo rc=$?
kill $MP
sort -n gpurun_out/c5_mem.txt | tail -1
cat gpurun_out/bench_c5_n1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','iters_per_s','itn_per_step','frac_of_hbm_roofline')}); print(d['config']); print(d['e2e']); print(d['roofline'])"
tail -5 gpurun_out/bench_c5_n1.err

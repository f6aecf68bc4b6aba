echo "== cfg2 x 40 solves (parked blocks)"; SVDLAS2_TRACE=1 timeout 600 python scripts/solve_profile.py --workload cfg2 --reps 40 > gpurun_out/r2ac_trace_cfg2.log 2>&1; grep -E "^\{" gpurun_out/r2ac_trace_cfg2.log | python -c "
import sys,json
ts=[json.loads(l)['t_total'] for l in sys.stdin]
print('t_total sorted', sorted(round(t,3) for t in ts))"
grep "device allocator" gpurun_out/r2ac_trace_cfg2.log | sort -t'(' -k2 | tail -3
echo "== bench cfg2 e2e"; timeout 600 python bench.py --workload cfg2 --steps 25 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2ac_bench_cfg2.json 2> gpurun_out/r2ac_bench_cfg2.err; python -c "
import json; d=json.load(open('gpurun_out/r2ac_bench_cfg2.json')); print(d['ms_per_step'], d['e2e'])"
echo "== cfg5 x 5"; SVDLAS2_TRACE=1 timeout 600 python scripts/solve_profile.py --workload cfg5 --reps 5 > gpurun_out/r2ac_trace_cfg5.log 2>&1; grep -E "^\{" gpurun_out/r2ac_trace_cfg5.log | cut -c1-300; grep "device allocator" gpurun_out/r2ac_trace_cfg5.log

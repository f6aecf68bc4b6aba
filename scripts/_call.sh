echo "== cfg2 x 40 solves"; SVDLAS2_TRACE=1 timeout 600 python scripts/solve_profile.py --workload cfg2 --reps 40 > gpurun_out/r2ab_trace_cfg2.log 2>&1; grep -E "^cgroup" gpurun_out/r2ab_trace_cfg2.log; grep -E "^\{" gpurun_out/r2ab_trace_cfg2.log | python -c "
import sys,json
ts=[json.loads(l)['t_total'] for l in sys.stdin]
print('t_total sorted', sorted(round(t,3) for t in ts))"
grep -E "throttle delta" gpurun_out/r2ab_trace_cfg2.log | sort | uniq -c | sort -rn | head -8

for w in cfg2 cfg5; do echo "== $w ritvec variance"; SVDLAS2_TRACE=1 timeout 300 python scripts/solve_profile.py --workload $w --reps 6 > gpurun_out/r2z_trace_$w.log 2>&1; grep -E "^\{" gpurun_out/r2z_trace_$w.log | cut -c1-330; done
bash scripts/ncu_r2v.sh > gpurun_out/r2v_ncu_call.log 2>&1; tail -12 gpurun_out/r2v_ncu_call.log

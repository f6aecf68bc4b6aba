#!/bin/bash
# Round-2 (second session) Nsight Compute evidence for the kernels added after scripts/ncu_r2.sh: the pipelined QL replay, its
# one-warp predecessor, and the whole-step kernel of small operators.  The target command (cfg 1 solve) has exited 0 on a B200
# without ncu first (gpurun_out/r2x_call.log, r2y_call.log).  Raw pages are exported next to the reports.
set -x
OUT=gpurun_out
NCU="timeout 240 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/r2v_launches_cfg1.csv python scripts/solve_profile.py --workload cfg1 --reps 1 > $OUT/r2v_ncu_launches.log 2>&1
$NCU --set full --import-source on -k regex:k_ql_replay_pipe -c 1 -o $OUT/r2v_ql_replay_pipe_cfg1 -f python scripts/solve_profile.py --workload cfg1 --reps 1 > $OUT/r2v_ncu_ql_pipe.log 2>&1
SVDLAS2_QL_PIPE=0 $NCU --set full --import-source on -k regex:k_ql_replay_smem -c 1 -o $OUT/r2v_ql_replay_one_warp_cfg1 -f python scripts/solve_profile.py --workload cfg1 --reps 1 > $OUT/r2v_ncu_ql_one.log 2>&1
$NCU --set full --import-source on -k regex:k_step_chain -s 200 -c 1 -o $OUT/r2v_step_fused_cfg1 -f python scripts/solve_profile.py --workload cfg1 --reps 1 > $OUT/r2v_ncu_step.log 2>&1
for r in r2v_ql_replay_pipe_cfg1 r2v_ql_replay_one_warp_cfg1 r2v_step_fused_cfg1; do
  [ -f $OUT/$r.ncu-rep ] && ncu -i $OUT/$r.ncu-rep --page raw --csv > $OUT/$r.raw.csv 2>/dev/null
done
ls -la $OUT/r2v_*

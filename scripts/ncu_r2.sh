#!/bin/bash
# Round-2 Nsight Compute evidence, one gpurun call (all ncu runs of a call count as one).  Every target command has already
# exited 0 on a B200 without ncu (gpurun_out/r2f_call.log, r2b_call.log); gpurun repeats that plain run itself before profiling.
#   launch list of one cfg4 solve  ->  gpurun_out/r2_launches_cfg4.csv
#   --set full captures            ->  gpurun_out/r2_*.ncu-rep  (+ the raw-page CSVs next to them)
set -x
OUT=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 9500 --csv --log-file $OUT/r2_launches_cfg4.csv python scripts/solve_profile.py --workload cfg4 --reps 1 > $OUT/r2_ncu_launches.log 2>&1
# the three SpMV launches of one opb at cfg4 (B: hot prefix; B': heavy rows slab-local, light rows wide slabs) + their reduce kernels
$NCU --set full --import-source on -k regex:"k_spmv_chunk|k_slab_reduce" -s 15 -c 5 -o $OUT/r2_spmv_cfg4 -f python scripts/profile_opb.py --workload cfg4 --reps 1 > $OUT/r2_ncu_spmv.log 2>&1
# everything else, one kernel family per pass, on a full cfg2 solve (upload once per pass: ~10 s each)
for k in k_axpy_dot:40:2 k_panel_dots:4:2 k_panel_update:4:2 k_ql_replay:0:1 k_spmm4_chunk:2:2 k_unpack4_sumsq:2:1 k_scale:40:1 k_chunk_fixup:40:1 k_allreduce_epilogue:0:0; do
  name=${k%%:*}; rest=${k#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  [ "$cnt" = "0" ] && continue
  $NCU --set full --import-source on -k regex:$name -s $skip -c $cnt -o $OUT/r2_${name}_cfg2 -f python scripts/solve_profile.py --workload cfg2 --reps 1 > $OUT/r2_ncu_$name.log 2>&1
done
# both Ritz contraction kernels at the cfg5 shape (js 2252, d 1000)
$NCU --set full --import-source on -k regex:k_ritz_contract -s 1 -c 1 -o $OUT/r2_ritz_fma_cfg5shape -f python scripts/ritz_times.py 500000 2252 1000 1 > $OUT/r2_ncu_ritz_fma.log 2>&1
$NCU --set full --import-source on -k regex:k_ritz_contract_dmma -s 1 -c 1 -o $OUT/r2_ritz_dmma_cfg5shape -f python scripts/ritz_times.py 500000 2252 1000 1 > $OUT/r2_ncu_ritz_dmma.log 2>&1
ls -la $OUT/r2_*.ncu-rep $OUT/r2_launches_cfg4.csv

#!/bin/bash
# ncu --set full captures of the kernels outside the headline rollout (one gpurun call)
o=gpurun_out
run() {  # name, kernel regex, skip, count, command...
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  "$@" > $o/ncu_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c $cnt -f -o $o/r02_$name "$@" > $o/ncu_$name.log 2>&1
  tail -1 $o/ncu_$name.log | cut -c1-200
}
run replay  'k_push|k_tree|k_sample|k_gather' 6 12 python tuning/prof_kernels.py replay
run substep 'k_encode|k_bfs_action|k_step_turn|k_begin_tick|k_end_tick' 330 5 python tuning/prof_kernels.py substep
run bfs     'k_bfs_batch|k_bfs_field' 2 2 python tuning/prof_kernels.py bfs
run shaped  'k_tick_lanes' 90 1 python tuning/prof_kernels.py shaped
run c5      '^k_rollout$' 12 1 python bench.py --workload c5 --steps 3 --warmup 3 --reps 1 --ticks 8 --no-cpu-baseline --no-sub --spinup 16
run c3      '^k_rollout$' 12 1 python bench.py --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-sub --spinup 8

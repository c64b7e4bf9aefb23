#!/bin/bash
# ncu --set full captures of the kernels outside the headline rollout + the rollout at c3 / c5 (one gpurun call).
# Only the raw-page CSVs travel back (gpurun_out is limited to 64 MiB); the c3 report itself is kept.
o=gpurun_out
run() {  # name, kernel regex, skip, count, command...
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  "$@" > $o/ncu_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c $cnt -f -o $o/r02_$name "$@" > $o/ncu_$name.log 2>&1
  tail -1 $o/ncu_$name.log | cut -c1-200
  ncu -i $o/r02_$name.ncu-rep --page raw --csv > $o/r02_$name.raw.csv 2>/dev/null
  if [ "$name" != "c3" ]; then rm -f $o/r02_$name.ncu-rep; fi
}
run replay  'k_push|k_tree|k_sample|k_gather' 6 12 python tuning/prof_kernels.py replay
run substep 'k_encode|k_bfs_action|k_step_turn|k_begin_tick|k_end_tick' 330 5 python tuning/prof_kernels.py substep
run bfs     'k_bfs_batch|k_bfs_field' 2 2 python tuning/prof_kernels.py bfs
run shaped  'k_tick_lanes' 90 1 python tuning/prof_kernels.py shaped
run c5      '^k_rollout$' 12 1 python bench.py --workload c5 --steps 3 --warmup 3 --reps 1 --ticks 8 --no-cpu-baseline --no-sub --spinup 16
run c3      '^k_rollout$' 30 1 python bench.py --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-sub
ncu -i $o/r02_c3.ncu-rep --page source --csv > $o/r02_c3.source.csv 2>/dev/null
# launch list of the default bench command (kernel shares of the step)
python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-sub > $o/launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $o/r02_launches_c3.csv python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-sub > $o/launch_ncu.log 2>&1
ls -la $o | head -40

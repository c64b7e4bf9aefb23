#!/bin/bash
# Expert collection (c5, A_star mode + next_obs + replay push) against BFS workers per CTA / worker poll interval
for lib in - nb2 nb3 slowpoll -; do
  if [ "$lib" != "-" ]; then export AGV_B200_LIB=$PWD/tuning/lib_$lib.so; else unset AGV_B200_LIB; fi
  echo "lib=$lib $(timeout 300 python bench.py --mode expert --steps 4 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"])')"
done
for lib in - nb2 nb3; do
  if [ "$lib" != "-" ]; then export AGV_B200_LIB=$PWD/tuning/lib_$lib.so; else unset AGV_B200_LIB; fi
  timeout 300 python bench.py --workload c5 --ticks 8 --policy astar --steps 6 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>/dev/null | python tuning/brief.py "lib=$lib c5 astar"
done

#!/bin/bash
# c5 (128x128): BFS workers per CTA x CTAs per SM, for the guided rollout, A_star mode and Expert collection
run() {  # lib cps
  if [ "$1" != "-" ]; then export AGV_B200_LIB=$PWD/tuning/lib_$1.so; else unset AGV_B200_LIB; fi
  export AGV_ROLLOUT_CPS=$2
  g=$(timeout 300 python bench.py --workload c5 --ticks 8 --steps 6 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.3e"%d["value"])')
  a=$(timeout 300 python bench.py --workload c5 --ticks 8 --policy astar --steps 6 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.3e"%d["value"])')
  x=$(timeout 300 python bench.py --mode expert --steps 4 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.3e"%d["value"])')
  echo "lib=$1 cps=$2 guided=$g astar=$a expert=$x"
}
run nb2 2; run nb3 2; run nb2 1; run nb3 1; run - 1; run nb5 1; run nb6 1; run nb8 1

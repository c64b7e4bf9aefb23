#!/bin/bash
# A/B of the in-tree library against tuning/libagv_prev.so on the same box
for args in "--obs clip" "--obs full" "--policy astar"; do
  for lib in prev cur; do
    if [ $lib = prev ]; then export AGV_B200_LIB=$PWD/tuning/libagv_prev.so; else unset AGV_B200_LIB; fi
    python bench.py $args --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py $lib $args
  done
done

#!/bin/bash
# A/B of library variants on the rollout bench:  bash tuning/r2_ab.sh "<bench args>" lib1 lib2 ...  ("-" = in-tree)
args=$1; shift
for lib in "$@"; do
  if [ "$lib" != "-" ]; then export AGV_B200_LIB=$PWD/tuning/lib_$lib.so; else unset AGV_B200_LIB; fi
  timeout 300 python bench.py $args --steps 12 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>/dev/null | python tuning/brief.py "lib=$lib" $args
done

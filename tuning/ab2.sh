#!/bin/bash
for rep in 1 2; do
for lib in prev cur v1 v2; do
  if [ $lib = cur ]; then unset AGV_B200_LIB; else export AGV_B200_LIB=$PWD/tuning/libagv_$lib.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py $lib
done
done

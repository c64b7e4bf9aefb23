#!/bin/bash
for lib in cur mb8_3 mb8_4; do
  if [ $lib = cur ]; then unset AGV_B200_LIB; else export AGV_B200_LIB=$PWD/tuning/libagv_$lib.so; fi
  python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py c5 $lib
done

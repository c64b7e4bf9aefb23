#!/bin/bash
for rep in 1 2; do
for lib in cur u2 u8; do
  if [ $lib = cur ]; then unset AGV_B200_LIB; else export AGV_B200_LIB=$PWD/tuning/libagv_$lib.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py $lib
done
done
for w in 2 4; do AGV_WARPS_PER_CTA=$w python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py cur wpc=$w; done

#!/bin/bash
# A/B of library builds on the same box (run under gpurun):
#   bash tuning/ab.sh "<bench args>" [lib.so ...]     "" = the in-tree library
# Build a variant with  AGV_NVCC_EXTRA="-D..." python <pkg>/build.py  and copy the .so to tuning/.
# AGV_TICK_IMPL=async|lanes|warp selects the fused-tick implementation inside one library.
args=$1; shift
for lib in "" "$@"; do
  if [ -n "$lib" ]; then export AGV_B200_LIB=$PWD/$lib; else unset AGV_B200_LIB; fi
  for rep in 1 2; do
    python bench.py $args --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python tuning/brief.py "lib=${lib:-in-tree}" $args
  done
done

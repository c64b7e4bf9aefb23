#!/bin/bash
# ncu --set full capture of one k_tick launch of the c3 bench + per-function attribution.
#   gpurun: bash tuning/prof.sh <name> [extra bench args]     -> gpurun_out/<name>.ncu-rep
name=$1; shift
ncu --set full --clock-control none --import-source on -k regex:k_tick -s 132 -c 1 -f -o gpurun_out/$name \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log | cut -c1-200

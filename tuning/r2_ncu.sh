#!/bin/bash
# ncu --set full capture of one k_rollout launch (c3, K=16) with source-level stall sampling
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --rollout 16 --spinup 8"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^k_rollout$ -s 12 -c 1 -f -o gpurun_out/${1:-r02_rollout} $CMD > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log | cut -c1-300

#!/bin/bash
# first GPU pass of the rollout kernel: parity tests, then a few bench points
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rollout.py -x -q 2>&1 | tail -25 > gpurun_out/r2_test_rollout.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rollout or fuzz" 2>&1 | tail -15 > gpurun_out/r2_test_parity_rollout.log
for cfg in "0 0" "8 0" "16 0" "16 3" "16 2" "32 2" "16 1"; do
  set -- $cfg
  AGV_ROLLOUT_CPS=$2 timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --rollout $1 2>gpurun_out/r2_bench_$1_$2.err | tail -1 > gpurun_out/r2_bench_$1_$2.json
done

#!/bin/bash
# compute-sanitizer memcheck over the rollout parity tests (one tool per call)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q -k "c3_guided_clip or c3_astar_clip or wide_128 or both_launch_shapes or c3_given or n7_not or tall_40" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -15 gpurun_out/r02_sanitizer_memcheck.log

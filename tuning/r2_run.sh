#!/bin/bash
# rollout parity tests, bench A/B over library variants, phase profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rollout.py -x -q 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rollout or fuzz" 2>&1 | tail -5
bash tuning/r2_ab.sh "--rollout 16" - $VARIANTS
CPS_LIST=4 bash tuning/r2_prof.sh

#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -2
bash tuning/r2_ab.sh "" spec0 spec1 spec0 spec1
bash tuning/r2_ab.sh "--policy astar" spec0 spec1

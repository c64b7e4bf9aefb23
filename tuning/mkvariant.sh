#!/bin/bash
# builds a library variant with extra nvcc flags as tuning/lib_<name>.so (CPU box), leaving the in-tree
# default library alone:   bash tuning/mkvariant.sh <name> "<flags>"
# on the GPU box:          AGV_B200_LIB=$PWD/tuning/lib_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
P=reinforcement-learning-for-multi-agv-pathfinding_b200
AGV_NVCC_EXTRA="$2" AGV_BUILD_OUT=$PWD/tuning/lib_$1.so python $P/build.py >/dev/null
echo tuning/lib_$1.so

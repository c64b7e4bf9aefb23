#!/bin/bash
export AGV_B200_LIB=$PWD/tuning/lib_prof.so AGV_ROLL_PROF=1 AGV_PROF_DUMP=1
for k in ${K_LIST:-16}; do
  echo "=== prof K $k"; timeout 300 python tuning/roll_prof.py $k 2>&1 | grep -v "^agv prof: 0 0 0 0 0 0 " | python tuning/phases.py | cut -c1-260
done

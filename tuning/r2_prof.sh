#!/bin/bash
export AGV_B200_LIB=$PWD/tuning/lib_prof.so AGV_ROLL_PROF=1 AGV_PROF_DUMP=1
for cps in ${CPS_LIST:-4}; do
  echo "=== prof CPS $cps"; AGV_ROLLOUT_CPS=$cps timeout 300 python tuning/roll_prof.py 16 2>&1 | grep -A1 -E "PHASES|^K=|light scenes|scene duration|fit" | grep -v "^--" | python tuning/phases.py
done

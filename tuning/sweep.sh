#!/bin/bash
# tuning sweep: warps per CTA x register budget for k_tick (c3).  Run under gpurun.
for lib in "" tuning/libagv_minb8.so tuning/libagv_minb5.so; do
  for wpc in 1 2 4; do
    if [ -n "$lib" ]; then export AGV_B200_LIB=$PWD/$lib; else unset AGV_B200_LIB; fi
    export AGV_WARPS_PER_CTA=$wpc
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib=%s wpc=%s  value=%.4g ms=%.3f e2e=%.4g' % (os.environ.get('AGV_B200_LIB','default(minb6)').split('/')[-1], os.environ['AGV_WARPS_PER_CTA'], d['value'], d['ms_per_step'], d['e2e']['value']))"
  done
done

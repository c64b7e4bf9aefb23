#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -2
export AGV_PROF_DUMP=1
timeout 300 python bench.py --steps 10 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>gpurun_out/c3_g.err | python tuning/brief.py c3 guided; grep "rollout load" gpurun_out/c3_g.err | sort | uniq -c | sort -rn | head -3
timeout 300 python bench.py --policy astar --steps 6 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>gpurun_out/c3_a.err | python tuning/brief.py c3 astar; grep "rollout load" gpurun_out/c3_a.err | sort | uniq -c | sort -rn | head -3
timeout 300 python bench.py --given --steps 10 --warmup 3 --reps 3 --no-cpu-baseline --no-sub 2>/dev/null | python tuning/brief.py c3 given

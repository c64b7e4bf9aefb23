#!/bin/bash
# the round's measurement set (1 GPU); JSON lines land in gpurun_out/r02_bench_*.json
o=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $o/r02_gputests.log; cat $o/r02_gputests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $o/r02_bench_c3.json 2>$o/r02_bench_c3.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $o/r02_bench_reference.json 2>$o/r02_bench_reference.err
python bench.py --obs full --ticks 0 --no-cpu-baseline --no-sub > $o/r02_bench_c3_full.json 2>/dev/null
python bench.py --policy astar --no-cpu-baseline --no-sub > $o/r02_bench_c3_astar.json 2>/dev/null
python bench.py --shaped --ticks 4 --steps 5 --no-cpu-baseline --no-sub > $o/r02_bench_c3_shaped.json 2>/dev/null
python bench.py --ticks 0 --no-cpu-baseline --no-sub > $o/r02_bench_c3_singletick.json 2>/dev/null
python bench.py --workload c2 --ticks 32 --no-sub > $o/r02_bench_c2.json 2>/dev/null
python bench.py --workload c5 --ticks 8 --steps 10 --no-sub > $o/r02_bench_c5.json 2>/dev/null
python bench.py --mode substep --no-cpu-baseline > $o/r02_bench_c3_substep.json 2>/dev/null
for f in $o/r02_bench_*.json; do echo "== $f"; python tuning/brief.py < $f 2>/dev/null || tail -c 300 $f; done

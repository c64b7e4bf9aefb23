#!/bin/bash
# the round's final measurement set (1 GPU); JSON lines land in gpurun_out/final_*.json
o=gpurun_out
python bench.py > $o/final_c3.json 2>$o/final_c3.err
python bench.py --obs full --no-cpu-baseline > $o/final_c3_full.json 2>/dev/null
python bench.py --policy astar --no-cpu-baseline > $o/final_c3_astar.json 2>/dev/null
python bench.py --shaped --no-cpu-baseline > $o/final_c3_shaped.json 2>/dev/null
python bench.py --workload c2 > $o/final_c2.json 2>/dev/null
python bench.py --workload c5 --steps 20 > $o/final_c5.json 2>/dev/null
python bench.py --mode substep --no-cpu-baseline > $o/final_c3_substep.json 2>/dev/null
python bench.py --mode expert --steps 20 --no-cpu-baseline > $o/final_c3_expert.json 2>/dev/null
python bench.py --mode expert --workload c5 --steps 10 --no-cpu-baseline > $o/final_c5_expert.json 2>/dev/null
python bench.py --mode train --workload c2 --steps 20 --no-cpu-baseline > $o/final_c2_train.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $o/final_ref.json 2>/dev/null
for f in $o/final_*.json; do echo "== $f"; python tuning/brief.py < $f 2>/dev/null || tail -c 400 $f; done

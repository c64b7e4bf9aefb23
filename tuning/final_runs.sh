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
# refresh of the headline kernel's ncu capture + launch list
run() {
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  "$@" > $o/ncu_plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c $cnt -f -o $o/r02_$name "$@" > $o/ncu_$name.log 2>&1
  ncu -i $o/r02_$name.ncu-rep --page raw --csv > $o/r02_$name.raw.csv 2>/dev/null
}
run c3 '^k_rollout$' 30 1 python bench.py --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-sub
run c5 '^k_rollout$' 52 1 python bench.py --workload c5 --steps 3 --warmup 3 --reps 1 --ticks 8 --no-cpu-baseline --no-sub
rm -f $o/r02_c5.ncu-rep
python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-sub > $o/launch_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $o/r02_launches_c3.csv python bench.py --steps 4 --warmup 3 --reps 1 --no-cpu-baseline --no-sub > $o/launch_ncu.log 2>&1

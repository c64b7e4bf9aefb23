#!/bin/bash
# (CPU box) per-function instruction attribution of gpurun_out/<name>.ncu-rep for kernel symbol regex $2 in cubin $3
name=$1; sym=$2; cub=${3:-agv_lanes}; turns=${4:-493000}
d=$(mktemp -d); ( cd $d && cuobjdump -xelf all /root/repo/reinforcement-learning-for-multi-agv-pathfinding_b200/libagv_b200.so >/dev/null )
idx=$(cuobjdump -elf $d/$cub.sm_100a.cubin 2>/dev/null | grep -E "^\s*0x[0-9a-f]+ .*0x12 .* $sym" | awk '{print $1}' | head -1)
nvdisasm -g -fun $idx $d/$cub.sm_100a.cubin > $d/k.sass
ncu -i gpurun_out/$name.ncu-rep --page source --csv > $d/src.csv 2>/dev/null
python tuning/attribute.py $d/src.csv $d/k.sass ${sym#_ZN3agv*[0-9]} $turns
echo $d

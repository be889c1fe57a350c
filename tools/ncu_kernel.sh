#!/bin/bash
# usage: tools/ncu_kernel.sh <kernel regex> <out name> [reads]   (run under gpurun, one GPU)
K=$1; O=$2; N=${3:-5000000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/$O -f \
    python bench.py --reads $N --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/$O.log 2>&1
ncu -i gpurun_out/$O.ncu-rep --page raw --csv > gpurun_out/$O.raw.csv 2>/dev/null
tail -2 gpurun_out/$O.log | cut -c1-300

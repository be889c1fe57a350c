#!/bin/bash
# per-launch device time + DRAM bytes of one bench step (run under gpurun, one GPU); $1 = output stem, $2 = reads
O=${1:-launches}; N=${2:-50000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/$O.csv python bench.py --reads $N --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/$O.log 2>&1
python tools/launch_summary.py gpurun_out/$O.csv > gpurun_out/$O.txt

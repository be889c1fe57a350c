#!/bin/bash
# round-2 evidence for the headline step (one B200, 50M x 150 bp, 2048 cores): per-launch device time + DRAM bytes of one whole step,
# and one `ncu --set full` capture of the top kernels. Numbers taken under ncu are never bench values.
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2/launches.csv > gpurun_out/r2/launches.txt; tail -45 gpurun_out/r2/launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'resolve_dense_k|scan_smem2_k|gather_rows16_k|emit_reads_fast_k|emit_names_st_k|sort_scatter_k|emit_off_reduce_k|build_keys_pk_k' -c 16 -o gpurun_out/r2/top -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/top.log 2>&1
ncu -i gpurun_out/r2/top.ncu-rep --page raw --csv > gpurun_out/r2/top.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2/top.raw.csv > gpurun_out/r2/top.txt 2>/dev/null; grep -c "^==" gpurun_out/r2/top.txt

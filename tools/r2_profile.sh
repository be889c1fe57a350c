#!/bin/bash
# round-2 evidence for the headline step (one B200, 50M x 150 bp, 2048 cores): one `ncu --set full` capture of the top kernels
# (the launch list of the same command is taken by tools/r2_final.sh). Numbers taken under ncu are never bench values.
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'resolve_dense_k|scan_smem2_k|gather_rows16_k|emit_reads_fast_k|emit_names_st_k|sort_scatter_k|emit_off_reduce_k|build_keys_pk_k' -c 16 -o gpurun_out/r2/top -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/top.log 2>&1
ncu -i gpurun_out/r2/top.ncu-rep --page raw --csv > gpurun_out/r2/top.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2/top.raw.csv > gpurun_out/r2/top.txt 2>/dev/null; grep -c "^==" gpurun_out/r2/top.txt
rm -f gpurun_out/r2/top.ncu-rep

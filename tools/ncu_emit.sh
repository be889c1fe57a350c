#!/bin/bash
# full ncu capture of one launch of each output-side kernel (5M reads)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gather_rows16_k|emit_reads_st_k|emit_names_st_k|resolve_finalize_k' -c 4 -o gpurun_out/emit_v2 -f \
    python bench.py --reads 5000000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/emit_v2.log 2>&1
ncu -i gpurun_out/emit_v2.ncu-rep --page raw --csv > gpurun_out/emit_v2.raw.csv 2>/dev/null

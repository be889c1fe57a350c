#!/bin/bash
# One-GPU evidence for profiles/: full bench line, reference arm, ncu launch list, ncu --set full of the top kernels.
# Run under gpurun (one GPU); copies land in gpurun_out/ and are summarised into profiles/ by tools/collect_profiles.py.
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
tools/launch_list.sh ${R}_launches 50000000
ncu --set full --clock-control none --import-source on \
    -k regex:'resolve_dense_k|scan_smem_k|gather_rows16_k|emit_reads_st_k|emit_names_st_k|gather_meta_k|build_keys_pk_k|tie_mid_groups_k' -c 8 \
    -o gpurun_out/${R}_top -f python bench.py --reads 5000000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${R}_top.log 2>&1
ncu -i gpurun_out/${R}_top.ncu-rep --page raw --csv > gpurun_out/${R}_top.raw.csv 2>/dev/null
tail -c 600 gpurun_out/${R}_bench_n1.json
# the radix-sort kernels separately (one pass is representative; they would otherwise use up the launch count)
ncu --set full --clock-control none --import-source on -k regex:'sort_scatter_k|sort_hist_k' -c 2 \
    -o gpurun_out/${R}_sort -f python bench.py --reads 5000000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${R}_sort.log 2>&1
ncu -i gpurun_out/${R}_sort.ncu-rep --page raw --csv > gpurun_out/${R}_sort.raw.csv 2>/dev/null

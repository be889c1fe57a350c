#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_bigcore.py tests/test_gpu_dropin.py tests/test_gpu_container.py -q > gpurun_out/r2/tests6.log 2>&1; tail -25 gpurun_out/r2/tests6.log
timeout 600 python bench.py --cores 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/c2_1Mcores_b.json 2> gpurun_out/r2/c2_1Mcores_b.err; python tools/bench_brief.py gpurun_out/r2/c2_1Mcores_b.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_big_k|sp_flags_k|sp_counts_k|sp_decide_k' -c 7 -o gpurun_out/r2/big_kernels -f \
    python bench.py --cores 1000000 --reads 5000000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2/big_kernels.log 2>&1
ncu -i gpurun_out/r2/big_kernels.ncu-rep --page raw --csv > gpurun_out/r2/big_kernels.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2/big_kernels.raw.csv 2>/dev/null | head -120

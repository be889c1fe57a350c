#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_bigcore.py tests/test_gpu_container.py -q -x > gpurun_out/r2/tests7.log 2>&1; tail -5 gpurun_out/r2/tests7.log
for q in 32 16; do
  SCB_BIG_Q=$q timeout 300 python bench.py --cores 1000000 --reads 10000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/bigq_$q.json 2> gpurun_out/r2/bigq_$q.err
  echo "== Q $q rc=$?"; python tools/bench_brief.py gpurun_out/r2/bigq_$q.json | head -2
done
timeout 600 python bench.py --cores 1000000 --steps 2 --warmup 1 --no-e2e --cpu-sample 100000 > gpurun_out/r2/c2_1Mcores_c.json 2> gpurun_out/r2/c2_1Mcores_c.err; python tools/bench_brief.py gpurun_out/r2/c2_1Mcores_c.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_big_k|pack_reads16_k' -c 2 -o gpurun_out/r2/big_kernels2 -f \
    python bench.py --cores 1000000 --reads 5000000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2/big_kernels2.log 2>&1
ncu -i gpurun_out/r2/big_kernels2.ncu-rep --page raw --csv > gpurun_out/r2/big_kernels2.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2/big_kernels2.raw.csv 2>/dev/null | head -70

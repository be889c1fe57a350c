#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2/tests11.log 2>&1; tail -6 gpurun_out/r2/tests11.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/c2_clean.json 2> gpurun_out/r2/c2_clean.err; python tools/bench_brief.py gpurun_out/r2/c2_clean.json | head -2
for wm in 1 0; do
SCB_SPARSE_WARM=$wm SCB_SPARSE_PROF=1 timeout 600 python bench.py --cores 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/c2_1Mcores_w$wm.json 2> gpurun_out/r2/c2_1Mcores_w$wm.err; python tools/bench_brief.py gpurun_out/r2/c2_1Mcores_w$wm.json | head -2
grep "sparse round" gpurun_out/r2/c2_1Mcores_w$wm.err | tail -34 | awk 'NR%3==1'
done

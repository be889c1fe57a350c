#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_container.py -q -x > gpurun_out/r2/tests9.log 2>&1; tail -6 gpurun_out/r2/tests9.log
SCB_SPARSE_PROF=1 timeout 600 python bench.py --cores 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/c2_1Mcores_e.json 2> gpurun_out/r2/c2_1Mcores_e.err; python tools/bench_brief.py gpurun_out/r2/c2_1Mcores_e.json | head -3
grep "sparse round" gpurun_out/r2/c2_1Mcores_e.err | tail -32

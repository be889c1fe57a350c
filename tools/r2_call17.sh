#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_tool.py -q -k "streaming or host_tool" > gpurun_out/r2/tests17.log 2>&1; tail -25 gpurun_out/r2/tests17.log
timeout 900 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2/c4_n1.json 2> gpurun_out/r2/c4_n1.err; python tools/bench_brief.py gpurun_out/r2/c4_n1.json | head -3; tail -2 gpurun_out/r2/c4_n1.err | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_sharded.py -x -q -k "pair or config or line_cap or shapes or cpp" 2>&1 | tail -4
python bench.py --config c3 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/c3_reads2.json
python tools/bench_brief.py gpurun_out/r2/c3_reads2.json | grep -v clocks

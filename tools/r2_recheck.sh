#!/bin/bash
# after touching the emit gathers: the whole GPU suite and the kernel-only lines of c2 / c3 on one B200
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2/tests_final2.log; cat gpurun_out/r2/tests_final2.log
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/final2_c2.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r2/final2_c2.json | head -2
python bench.py --config c3 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/final2_c3.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r2/final2_c3.json | head -2

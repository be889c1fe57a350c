#!/bin/bash
# block-sequential sparse tie-break + long reads: parity tests, then the 1M-core headline with the default block schedule
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_bigcore.py tests/test_gpu_parity.py tests/test_gpu_sharded.py tests/test_gpu_container.py -x -q 2>&1 | tail -15
SCB_SPARSE_PROF=1 python bench.py --cores 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/sparse_blk_default.json 2> gpurun_out/r2/sparse_blk_default.err
python tools/bench_brief.py gpurun_out/r2/sparse_blk_default.json
grep "sparse block" gpurun_out/r2/sparse_blk_default.err | head -20
SCB_SPARSE_BLOCK=8388608 python bench.py --cores 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/sparse_blk_8M.json
python tools/bench_brief.py gpurun_out/r2/sparse_blk_8M.json
python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/c2_after_blk.json
python tools/bench_brief.py gpurun_out/r2/c2_after_blk.json

#!/bin/bash
# block-sequential sparse tie-break: parity tests, then the 1M-core headline at three block sizes
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_bigcore.py -x -q 2>&1 | tail -15
for B in 262144 1048576 4194304; do
  echo "== SCB_SPARSE_BLOCK=$B"
  SCB_SPARSE_BLOCK=$B SCB_SPARSE_PROF=1 python bench.py --cores 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/sparse_blk_$B.json 2> gpurun_out/r2/sparse_blk_$B.err
  python tools/bench_brief.py gpurun_out/r2/sparse_blk_$B.json
  grep "sparse block" gpurun_out/r2/sparse_blk_$B.err | tail -4
done

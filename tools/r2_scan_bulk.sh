#!/bin/bash
# scan tile staging: cp.async.bulk + mbarrier (default) against per-lane 16-byte cp.async (build/variants/lib_scan_ldgsts.so).
# The variant library is not kept in the tree; build it first (here, nvcc cross-compiles):
#   mkdir -p build/variants && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O2 -shared \
#     --expt-relaxed-constexpr -DSCB_SCAN_BULK=0 -o build/variants/lib_scan_ldgsts.so scalce_b200/csrc/api.cu scalce_b200/csrc/core_table.cpp
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q 2>&1 | tail -4
for rep in 1 2; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/scan_bulk_$rep.json
  SCB_LIB_PATH=$PWD/build/variants/lib_scan_ldgsts.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/scan_ldgsts_$rep.json
done
python tools/bench_brief.py gpurun_out/r2/scan_bulk_*.json gpurun_out/r2/scan_ldgsts_*.json | grep -v clocks
python bench.py --config c3 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/scan_bulk_c3.json
SCB_LIB_PATH=$PWD/build/variants/lib_scan_ldgsts.so python bench.py --config c3 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null > gpurun_out/r2/scan_ldgsts_c3.json
python tools/bench_brief.py gpurun_out/r2/scan_bulk_c3.json gpurun_out/r2/scan_ldgsts_c3.json | grep -v clocks

#!/bin/bash
mkdir -p gpurun_out/final
export PYTHONUNBUFFERED=1
SCB_TEST_EXPERIMENTAL=1 timeout 27 python -m pytest tests/experimental_cases.py -q -k "scan_v2_dense or scan_v2_million or emit_reads_v2_odd or all_single_gpu_variants_together" > gpurun_out/final/exp2.log 2>&1; echo "exp2 rc=$?"; tail -n 6 gpurun_out/final/exp2.log
SCB_EMIT_FUSED_SCAN=1 SCB_EMIT_READS_V2=1 SCB_SCAN_V2=1 timeout 11 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/final/three.json 2> gpurun_out/final/three.err; echo "three rc=$?"; cut -c1-260 gpurun_out/final/three.json

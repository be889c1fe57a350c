#!/bin/bash
# Last GPU seconds of the round (budget ~150 s): default-path parity first, then the new headline number, then a verdict on the
# single-GPU opt-in variants. Every step has its own time limit, so nothing is left running for gpurun's limit to kill.
mkdir -p gpurun_out/final
export PYTHONUNBUFFERED=1
timeout 35 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > gpurun_out/final/parity.log 2>&1; echo "parity rc=$?"; tail -n 3 gpurun_out/final/parity.log
timeout 40 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final/bench_default.json
SCB_TEST_EXPERIMENTAL=1 timeout 40 python -m pytest tests/experimental_cases.py -q -k "single_gpu_variants" > gpurun_out/final/exp.log 2>&1; echo "exp rc=$?"; tail -n 12 gpurun_out/final/exp.log
SCB_EMIT_FUSED_SCAN=1 SCB_EMIT_CORESIDENT=1 SCB_EMIT_READS_V2=1 SCB_SCAN_V2=1 SCB_OVERLAP_CHUNKS=1 SCB_RESOLVE_CHEAP_GUESS=1 timeout 30 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/final/bench_all_on.json 2> gpurun_out/final/bench_all_on.err; echo "bench_all rc=$?"; cut -c1-300 gpurun_out/final/bench_all_on.json

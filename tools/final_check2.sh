#!/bin/bash
# second (last) call of the round: one kernel-only bench line per single-GPU switch
mkdir -p gpurun_out/final
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; env "$@" timeout 13 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/final/$name.json 2> gpurun_out/final/$name.err; echo "$name rc=$?"; }
run reads_v2 SCB_EMIT_READS_V2=1
run fused_scan SCB_EMIT_FUSED_SCAN=1
run coresident SCB_EMIT_CORESIDENT=1
run cheap_guess SCB_RESOLVE_CHEAP_GUESS=1
run overlap_chunks SCB_OVERLAP_CHUNKS=1
run scan_v2 SCB_SCAN_V2=1
python tools/ab_summary.py gpurun_out/final

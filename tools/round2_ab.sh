#!/bin/bash
# ONE gpurun call (one GPU, ~6 min) that answers every open question the end of round 1 left behind:
#   1. do the opt-in variants pass parity (tests/experimental_cases.py)?
#   2. what does each switch do to the headline step (50M x 150 bp, stage times from bench.py)?
#   3. the pipelined end-to-end arm.
# Usage:  gpurun --timeout 900 -- 'bash tools/round2_ab.sh'      results: gpurun_out/ab/ + a table on stdout
mkdir -p gpurun_out/ab
export PYTHONUNBUFFERED=1
SCB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/experimental_cases.py -q > gpurun_out/ab/tests.log 2>&1
tail -n 15 gpurun_out/ab/tests.log
run() {  # name ENV=... : kernel-only bench line of one configuration
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab/$name.json 2> gpurun_out/ab/$name.err
}
run default
run sort_per_bucket SCB_SORT_PER_BUCKET=1
run names_v2 SCB_EMIT_NAMES_V2=1
# tie-break: deferred re-sweeps (model: 8.95 -> 7.7 ms), alone and with the faster-growing early blocks
run resolve_defer SCB_RESOLVE_DEFER=1
run resolve_defer_g8 SCB_RESOLVE_DEFER=1 SCB_RESOLVE_GROWTH=8 SCB_RESOLVE_SMALL=4194304
# block schedule of the tie-break (tools/sim_resolve.c: x4 growth up to 4M / 16M reads -> 105 / 99 rounds instead of 111, larger blocks)
run resolve_small_4M SCB_RESOLVE_SMALL=4194304
run resolve_small_16M SCB_RESOLVE_SMALL=16777216
# early growth x8 / x16 up to 4M reads: 87 / 80 rounds in the model (a round costs ~35 us even on a tiny block)
run resolve_g8_4M SCB_RESOLVE_GROWTH=8 SCB_RESOLVE_SMALL=4194304
run resolve_g16_4M SCB_RESOLVE_GROWTH=16 SCB_RESOLVE_SMALL=4194304
run best_guess SCB_EMIT_NAMES_V2=1 SCB_RESOLVE_DEFER=1 SCB_RESOLVE_GROWTH=8 SCB_RESOLVE_SMALL=4194304
# per-round phase times of the tie-break kernel (P / D / E / sync / scan per round, to stderr): what a round's fixed cost consists of
SCB_RESOLVE_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ab/resolve_prof.json 2> gpurun_out/ab/resolve_prof.err
grep "resolve totals\|resolve subtile" gpurun_out/ab/resolve_prof.err | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-depth 2 > gpurun_out/ab/e2e_depth2.json 2> gpurun_out/ab/e2e_depth2.err
python tools/ab_summary.py gpurun_out/ab

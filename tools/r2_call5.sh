#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
for hnt in 0 1 2 4 7; do
  SCB_BIG_HINTS=$hnt timeout 300 python bench.py --cores 1000000 --reads 10000000 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/hints_$hnt.json 2> gpurun_out/r2/hints_$hnt.err
  echo "== hints $hnt rc=$?"; python tools/bench_brief.py gpurun_out/r2/hints_$hnt.json | head -2; tail -1 gpurun_out/r2/hints_$hnt.err | cut -c1-250
done
SCB_BIG_HINTS=7 timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python bench.py --cores 1000000 --reads 500000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/r2/sanitizer.log 2>&1
grep -v "^$" gpurun_out/r2/sanitizer.log | head -40

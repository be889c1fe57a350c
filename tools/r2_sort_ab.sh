#!/bin/bash
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
for v in default sort8_4 sort12_4 sort8_5; do
  if [ $v = default ]; then unset SCB_LIB_PATH; else export SCB_LIB_PATH=$PWD/scalce_b200/libscalce_b200_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/sortab_$v.json 2> gpurun_out/r2/sortab_$v.err
  echo "== $v"; python tools/bench_brief.py gpurun_out/r2/sortab_$v.json | head -2
done
unset SCB_LIB_PATH
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 6 > gpurun_out/r2/e2e_locks.json 2> gpurun_out/r2/e2e_locks.err; python tools/bench_brief.py gpurun_out/r2/e2e_locks.json | head -4

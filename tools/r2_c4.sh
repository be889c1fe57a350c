#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 2 --warmup 1 --no-cpu --no-e2e "$@" > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json 2>/dev/null | head -4; tail -1 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
run c4 --config c4 --no-parity
if [ "$N" = "4" ]; then run c2 --steps 3 --warmup 2; fi

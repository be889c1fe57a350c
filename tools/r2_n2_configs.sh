#!/bin/bash
# N GPUs: the other named configs through the sharded path (quick lines, parity check once)
N=${1:-2}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json 2>/dev/null | head -4; tail -2 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
run c3 --config c3 --e2e-steps 2
run c5 --config c5 --no-e2e --no-parity
run c4 --config c4 --no-e2e --no-parity
run c2_1M --cores 1000000 --no-e2e

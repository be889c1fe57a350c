#!/bin/bash
# N GPUs (gpurun --gpus N): parity over NCCL + CUDA IPC (C++ orchestrator: bucket ranges, flush chunks), host<->device copy rates of
# all ranks at once with and without CPU binding, then the bench lines. Usage: gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/r2/n${N}_topo.txt 2>&1
for o in cpp_nccl cpp_nccl_chunks; do
  echo "== parity, $o"; SCB_ORCH=$o timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29504 tests/sharded_nccl_worker.py 120000 100 1048576 2>&1 | tail -2
done
for b in 0 1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29506 tools/pcie_multi.py --bind $b 2>/dev/null | grep '^{' > gpurun_out/r2/n${N}_pcie_bind$b.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r2/n${N}_pcie_bind$b.json"))
print("pcie bind=$b slowest rank GB/s", {k: round(v, 1) for k, v in d["slowest_rank_GBps"].items()}, "sum", {k: round(v, 1) for k, v in d["sum_GBps"].items()}, d["binding"][0])
PY
done
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 2 --no-cpu $EXTRA > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json; tail -2 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
EXTRA="--e2e-steps 3" run chunks
EXTRA="--no-e2e --no-parity" run buckets SCB_SHARD_SPLIT=buckets
if [ -n "$MORE" ]; then
  EXTRA="--no-e2e --no-parity --cores 1000000" run 1Mcores
  EXTRA="--no-e2e --no-parity --config c3" run c3
fi

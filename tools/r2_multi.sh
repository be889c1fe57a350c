#!/bin/bash
# N GPUs (gpurun --gpus N): parity over NCCL + CUDA IPC first (default path, then the in-kernel joint rounds + early emit), then the
# bench line per switch. Usage: gpurun --gpus 2 --timeout 1500 -- 'bash tools/r2_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== parity, default"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 tests/sharded_nccl_worker.py 120000 100 2097152 2>&1 | tail -3
echo "== parity, joint kernel + early emit"; SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29502 tests/sharded_nccl_worker.py 120000 100 2097152 2>&1 | tail -3
echo "== parity, C++ orchestrator + NCCL comm library"; SCB_ORCH=cpp_nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29504 tests/sharded_nccl_worker.py 120000 100 2097152 2>&1 | tail -3
echo "== parity, C++ orchestrator + torch collectives"; SCB_ORCH=cpp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29505 tests/sharded_nccl_worker.py 120000 100 2097152 2>&1 | tail -3
echo "== parity, sparse engine"; SCB_RESOLVE=sparse SCB_TABLE=global timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29503 tests/sharded_nccl_worker.py 120000 100 2097152 2>&1 | tail -3
run() {
  local name=$1; shift
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 2 --no-cpu $EXTRA > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json; tail -2 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
EXTRA="--e2e-steps 3" run default
EXTRA="--no-e2e --no-parity" run early_emit SCB_SHARD_EARLY_EMIT=1
EXTRA="--no-e2e --no-parity" run joint_kernel SCB_SHARD_JOINT_KERNEL=1
EXTRA="--no-e2e --no-parity" run joint_early SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1

#!/bin/bash
# The sharded run's opt-in switches at N GPUs (gpurun --gpus N): default, early emit, in-kernel joint rounds, both.
# (parity of the joint kernel first: SCB_TEST_EXPERIMENTAL=1 python -m pytest tests/experimental_cases.py -q -k joint_kernel)
# Usage:  gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_ab_multi.sh 8'
N=${1:-2}
mkdir -p gpurun_out/ab
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/ab/n${N}_$name.json 2> gpurun_out/ab/n${N}_$name.err
}
run default
run early_emit SCB_SHARD_EARLY_EMIT=1
run joint_kernel SCB_SHARD_JOINT_KERNEL=1
run joint_early SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1
run joint_defer SCB_SHARD_JOINT_KERNEL=1 SCB_RESOLVE_DEFER=1
run all_on SCB_SHARD_JOINT_KERNEL=1 SCB_RESOLVE_DEFER=1 SCB_SHARD_EARLY_EMIT=1
# per-round sweep statistics of the default joint rounds (full / incremental / redone subtiles per round and rank, to stderr)
SCB_RESOLVE_STAT=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ab/n${N}_stat.json 2> gpurun_out/ab/n${N}_stat.err
python tools/ab_summary.py gpurun_out/ab n${N}_

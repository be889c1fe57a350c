#!/bin/bash
# 8 GPUs (gpurun --gpus 8): parity over NCCL + CUDA IPC (default path, in-kernel joint rounds + early emit, C++ orchestrator), the
# headline line both ways, then the other named configs with whichever won.
N=${1:-8}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
par() { echo "== parity $1"; env "${@:2}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 tests/sharded_nccl_worker.py 400000 100 4194304 2>&1 | tail -2; }
par default SCB_X=0
par joint_early SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1
par cpp_nccl SCB_ORCH=cpp_nccl
run() {
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 2 --no-cpu "$@" > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json 2>/dev/null | head -4; tail -1 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
run c2_default --e2e-steps 3
SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1 run c2_joint_early --no-e2e --no-parity
SCB_SHARD_JOINT_KERNEL=1 run c2_joint --no-e2e --no-parity
best=$(python - <<'PY'
import json
def ms(p):
    try: return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])["ms_per_step"]
    except Exception: return 1e9
c = {"": ms("gpurun_out/r2/n8_c2_default.json"), "SCB_SHARD_JOINT_KERNEL=1 SCB_SHARD_EARLY_EMIT=1": ms("gpurun_out/r2/n8_c2_joint_early.json"), "SCB_SHARD_JOINT_KERNEL=1": ms("gpurun_out/r2/n8_c2_joint.json")}
print(min(c, key=c.get))
PY
)
echo "== best switches: '$best'"
export $best
run c3 --config c3 --no-e2e --no-parity
run c5 --config c5 --no-e2e --no-parity
run c4 --config c4 --no-e2e --no-parity
run c2_1M --cores 1000000 --no-e2e --no-parity --steps 2 --warmup 1

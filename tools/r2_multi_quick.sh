#!/bin/bash
# N GPUs: NCCL parity of the chunk-ownership flow, then the kernel-only bench line. gpurun --gpus N -- 'bash tools/r2_multi_quick.sh N'
N=${1:-2}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
echo "== parity, cpp_nccl_chunks"; SCB_ORCH=cpp_nccl_chunks timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29504 tests/sharded_nccl_worker.py 120000 100 1048576 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 4 --warmup 2 --no-cpu --no-e2e ${EXTRA:---no-parity} > gpurun_out/r2/n${N}_quick.json 2> gpurun_out/r2/n${N}_quick.err
echo "== rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_quick.json; tail -2 gpurun_out/r2/n${N}_quick.err | cut -c1-300

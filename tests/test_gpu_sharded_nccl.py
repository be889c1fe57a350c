"""The sharded path over real NCCL (one process per GPU). Needs >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("orch", ["python", "cpp_nccl", "cpp_nccl_chunks"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_over_nccl(world, orch):
    """orch = python: scalce_b200/shard.py over torch.distributed; cpp_nccl: scb_shard_flush (the C++ orchestrator) over
    libscalce_b200_nccl.so; cpp_nccl_chunks: the same without a merged stream, i.e. with whole flush chunks handed to the ranks
    instead of bucket ranges. Either way one process per GPU, NCCL + CUDA IPC peer stores, checked against the oracle; the worker
    leaves a record under gpurun_out/ (copies of the round's runs: profiles/sharded_nccl_w*.json)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "sharded_nccl_worker.py"), "60000", "100", str(1 << 21)]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, SCB_ORCH=orch), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_NCCL_OK" in r.stdout, r.stdout[-4000:]

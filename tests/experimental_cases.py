"""Opt-in variants of the sharded run that are off by default until an 8-GPU run decides (see tools/r2_multi.sh), and the C++ host
tool end to end. Not collected by the default run (no test_ prefix): tests/test_zz_gpu_experimental.py runs this file in a child
process when SCB_RUN_EXPERIMENTAL=1. Run directly with  SCB_TEST_EXPERIMENTAL=1 python -m pytest tests/experimental_cases.py -q

  SCB_SHARD_JOINT_KERNEL=1 all joint tie-break rounds inside one kernel per rank, histograms exchanged through peer memory
  SCB_SHARD_EARLY_EMIT=1   names / packed reads / meta records are emitted while the quality rows still travel

Same bar as everywhere else: bit-exact against the oracle.
"""
import os

import pytest

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("SCB_TEST_EXPERIMENTAL", "0") in ("", "0"),
                                                   reason="experimental variants: set SCB_TEST_EXPERIMENTAL=1")]


def _sharded(n, L, world, **kw):
    run_kw = {k: kw.pop(k) for k in list(kw) if k in ("use_names", "use_quals", "bucket_set_bytes", "bounds")}
    paired = kw.get("paired", False)
    cores, b, q1, q2, _ = util.make_case(n, L, **kw)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, **{k: v for k, v in run_kw.items() if k != "bounds"})
    ranks = util.run_sharded_loopback(cores, b, q1, q2, world, paired=paired, **run_kw)
    util.assert_sharded_same(o, ranks, paired=paired)
    return o, ranks


def test_early_emit_paired_multi_chunk(monkeypatch):
    monkeypatch.setenv("SCB_SHARD_EARLY_EMIT", "1")
    _sharded(20000, 100, 3, seed=154, paired=True, L2=75, bucket_set_bytes=1 << 20, bounds=[0, 1000, 13000, 20000])


def test_early_emit_no_names_short(monkeypatch):
    monkeypatch.setenv("SCB_SHARD_EARLY_EMIT", "1")
    _sharded(40000, 36, 8, seed=155, use_names=False)


def test_early_emit_four_ranks(monkeypatch):
    monkeypatch.setenv("SCB_SHARD_EARLY_EMIT", "1")
    _sharded(30000, 100, 4, seed=157)


# ---- the C++ host tool end to end (scalce_b200/host/scb_boost.cpp): FASTQ -> temp files == the oracle's chunk streams ----
@pytest.mark.parametrize("world", [2, 4, 8])
def test_joint_kernel_over_nvlink(world):
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(root, "tests", "sharded_nccl_worker.py"), "120000", "100", str(1 << 21)]
    env = dict(os.environ, SCB_SHARD_JOINT_KERNEL="1", SCB_SHARD_EARLY_EMIT="1")
    r = subprocess.run(cmd, cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_NCCL_OK" in r.stdout, r.stdout[-4000:]

"""torchrun worker: the sharded path over NCCL, checked on rank 0 against the oracle over the whole input.
Launched by tests/test_gpu_sharded_nccl.py (and usable by hand under gpurun --gpus N)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from tests import util


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    bsb = int(sys.argv[3]) if len(sys.argv) > 3 else (1 << 21)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from scalce_b200.binding import BoostTransform
    from scalce_b200.shard import ShardedTransform, TorchComm, shard_bounds
    cores, b, q1, q2, _ = util.make_case(n, L, seed=77)
    bd = shard_bounds(n, world)
    a, z = bd[rank], bd[rank + 1]
    orch = os.environ.get("SCB_ORCH", "python")
    merged = orch != "cpp_nccl_chunks"      # no merged stream: scb_shard_flush hands out whole flush chunks instead of bucket ranges
    t = BoostTransform(cores, L, 0, bucket_set_bytes=bsb, device=local, emit_merged=merged)
    t.submit(b.seq[a:z], q1[a:z], b.names, b.name_off[a:z + 1])
    if orch == "cpp_nccl_chunks":
        os.environ["SCB_SHARD_SPLIT"] = "chunks"
    if orch in ("cpp_nccl", "cpp_nccl_chunks"):       # scb_shard_flush over libscalce_b200_nccl.so: the C++ orchestrator and C collectives, no Python on the path
        from scalce_b200.shard import CShardedTransform, NcclCComm
        st = CShardedTransform(t, NcclCComm(dist, local), use_torch_stream=False)
    elif orch == "cpp":          # scb_shard_flush with the collectives lent by torch.distributed through callbacks
        from scalce_b200.shard import CShardedTransform
        st = CShardedTransform(t, TorchComm(dist, torch.device("cuda", local)))
    else:
        st = ShardedTransform(t, TorchComm(dist, torch.device("cuda", local)))
    res = st.flush()
    mine = dict(dbg=res.debug(res.n_local), n_chunks=res.n_chunks, unb=t.unbucketed, stats=st.stats,
                streams={(k, c): res.stream(k, c) for k in range(4) for c in list(range(res.n_chunks)) + ([-1] if merged else [])})
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    ok = True
    if rank == 0:
        o = util.run_oracle(cores, b, q1, q2, bucket_set_bytes=bsb)
        dbg_o = o.debug()
        for k in ("node_id", "core", "end", "chunk"):
            got = np.concatenate([g["dbg"][k] for g in gathered])
            if not np.array_equal(dbg_o[k], got):
                print(f"MISMATCH per-read {k}"); ok = False
        for g in gathered:
            ok &= g["n_chunks"] == o.n_chunks
        if orch == "cpp_nccl_chunks":
            ok &= all(g["stats"].get("split") == "flush chunks" for g in gathered)
        for c in list(range(o.n_chunks)) + ([-1] if merged else []):
            for k in range(4):
                want = o.stream(k, c)
                got = b"".join(g["streams"][(k, c)] for g in gathered)
                if want != got:
                    print(f"MISMATCH chunk {c} stream {k}: {len(want)} vs {len(got)}"); ok = False
        ok &= o.unbucketed == sum(g["unb"] for g in gathered)
        print("rounds", gathered[0]["stats"]["rounds"], "chunks", o.n_chunks, "split", gathered[0]["stats"].get("split"))
        print("SHARDED_NCCL_OK" if ok else "SHARDED_NCCL_FAIL")
        # a record the driver / the round's profiles can keep: what ran, on how many GPUs, and the verdict
        import json
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        variant = "_".join(k + ("-" + os.environ[k] if k in ("SCB_RESOLVE", "SCB_ORCH") else "") for k in ("SCB_RESOLVE", "SCB_ORCH") if os.environ.get(k, "0") not in ("", "0")) or "default"
        rec = {"test": "tests/sharded_nccl_worker.py", "world_size": world, "gpus": torch.cuda.device_count(), "reads": n, "read_length": L,
               "bucket_set_bytes": bsb, "flush_chunks": o.n_chunks, "joint_rounds": gathered[0]["stats"]["rounds"], "variant": variant, "ownership": gathered[0]["stats"].get("split", "bucket ranges"),
               "backend": "nccl + CUDA IPC peer stores, one process per GPU", "ok": bool(ok),
               "compared": "per-read bucket / core / end / chunk arrays and streams 0-3 of every flush chunk" + (" and merged" if merged else "") + ", rank-order concatenation vs the CPU oracle"}
        with open(os.path.join(out_dir, f"sharded_nccl_w{world}_{variant}.json"), "w") as f:
            json.dump(rec, f)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

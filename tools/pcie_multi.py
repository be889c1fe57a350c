#!/usr/bin/env python3
"""All ranks of one node copy pinned host memory to and from their GPU at the same time: what the end-to-end arm of
bench.py can reach at N GPUs, with and without binding each rank to the CPUs next to its GPU (bench.bind_host_to_gpu).
torchrun --nproc-per-node N tools/pcie_multi.py [--bind 0|1] [--gib 2]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", type=int, default=1)
    ap.add_argument("--gib", type=float, default=2.0)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["SCB_BENCH_BIND"] = str(a.bind)
    binding = bench.bind_host_to_gpu(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(a.gib * (1 << 30))
    hin = torch.empty(n, dtype=torch.uint8).pin_memory()
    hin.fill_(1)                      # touch
    hout = torch.empty(n, dtype=torch.uint8).pin_memory()
    hout.fill_(2)
    dbuf = torch.empty(n, dtype=torch.uint8, device="cuda")
    dsrc = torch.ones(n, dtype=torch.uint8, device="cuda")
    res = {}
    s2 = torch.cuda.Stream()
    for what in ("h2d", "d2h", "both"):
        for rep in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if what in ("h2d", "both"):
                dbuf.copy_(hin, non_blocking=True)
            if what in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    hout.copy_(dsrc, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[what] = n / dt / 1e9
    t = torch.tensor([res["h2d"], res["d2h"], res["both"]], device="cuda", dtype=torch.float64)
    lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    allb = [None] * dist.get_world_size()
    dist.all_gather_object(allb, binding)
    if dist.get_rank() == 0:
        print(json.dumps({"bind": a.bind, "ranks": dist.get_world_size(), "gib_per_copy": a.gib,
                          "slowest_rank_GBps": {"h2d": float(lo[0]), "d2h": float(lo[1]), "both_per_direction": float(lo[2])},
                          "sum_GBps": {"h2d": float(sm[0]), "d2h": float(sm[1]), "both_per_direction": float(sm[2])},
                          "binding": allb}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Per-round subtile sweep statistics of the sharded tie-break (ranks as threads on one GPU). SCB_RESOLVE_STAT=1."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from scalce_b200 import synth
from scalce_b200.binding import BoostTransform
from scalce_b200.shard import LoopbackComm, ShardedTransform
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
per = int(sys.argv[2]) if len(sys.argv) > 2 else 6_000_000
cores = bench.headline_cores()
comms = LoopbackComm.make(world, 0)
def worker(r):
    d = synth.make_batch_cuda(per, 150, seed=1 + r, device="cuda:0")
    q = torch.where(d["seq"] == ord("N"), torch.zeros_like(d["qual"]), d["qual"] - 33)
    t = BoostTransform(cores, 150, device=0, emit_merged=False)
    t.submit_device(per, d["seq"].data_ptr(), q.data_ptr(), d["names"].data_ptr(), d["name_off"].data_ptr())
    st = ShardedTransform(t, comms[r])
    st.flush()
    if r == world - 1:
        print("rounds", st.stats["rounds"], file=sys.stderr)
th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
[x.start() for x in th]; [x.join() for x in th]

#!/bin/bash
# 8 GPUs (gpurun --gpus 8): topology, host<->device copy rates of all ranks at once with and without CPU binding, then the bench
# lines (the c2 line carries the NCCL parity checks of both ownership modes). Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_n8.sh'
N=${1:-8}
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/r2/n${N}_topo.txt 2>&1
nproc > gpurun_out/r2/n${N}_host.txt; lscpu | grep -i -E "numa|socket|model name" >> gpurun_out/r2/n${N}_host.txt
best=1
for b in 0 1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29506 tools/pcie_multi.py --bind $b 2>/dev/null | grep '^{' > gpurun_out/r2/n${N}_pcie_bind$b.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r2/n${N}_pcie_bind$b.json"))
print("pcie bind=$b slowest rank GB/s", {k: round(v, 1) for k, v in d["slowest_rank_GBps"].items()}, "sum", {k: round(v, 1) for k, v in d["sum_GBps"].items()})
print("   binding of ranks:", [(x.get("bound"), x.get("cpus"), x.get("numa_node"), x.get("first_cpu"), x.get("why")) for x in d["binding"]])
PY
done
best=$(python - <<PY
import json
a = json.load(open("gpurun_out/r2/n${N}_pcie_bind0.json"))["slowest_rank_GBps"]
b = json.load(open("gpurun_out/r2/n${N}_pcie_bind1.json"))["slowest_rank_GBps"]
print(1 if (b["h2d"] + b["d2h"]) >= 0.97 * (a["h2d"] + a["d2h"]) else 0)
PY
)
echo "binding for the bench: $best"
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 2 --no-cpu $EXTRA > gpurun_out/r2/n${N}_$name.json 2> gpurun_out/r2/n${N}_$name.err
  echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/n${N}_$name.json; tail -2 gpurun_out/r2/n${N}_$name.err | cut -c1-300
}
EXTRA="--e2e-steps 3" run c2 SCB_BENCH_BIND=$best
EXTRA="--no-e2e --no-parity --cores 1000000" run c2_1Mcores

#!/bin/bash
# one GPU: the headline line (with e2e + cpu baseline), the production-size core set, the other named configs
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2/$name.json 2> gpurun_out/r2/$name.err; echo "== $name rc=$?"; python tools/bench_brief.py gpurun_out/r2/$name.json; }
run c2_2048 --steps 5 --warmup 3
run c2_1Mcores --cores 1000000 --steps 2 --warmup 1 --no-e2e --cpu-sample 100000
run c3 --config c3 --steps 2 --warmup 1 --no-cpu --e2e-steps 2
run c5 --config c5 --steps 2 --warmup 1 --no-cpu --e2e-steps 2
run c4_250M --config c4 --reads 250000000 --steps 2 --warmup 1 --no-cpu --no-e2e
python tools/pcie_duplex.py 4 2>&1 | tee gpurun_out/r2/pcie.txt

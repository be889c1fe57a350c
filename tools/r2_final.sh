#!/bin/bash
# end-of-round evidence on one B200: the whole GPU suite, the headline line with e2e + CPU baseline, the other configs kernel-only,
# and the ncu launch list of the headline step (kernel shares). Numbers taken under ncu are never bench values.
mkdir -p gpurun_out/r2
export PYTHONUNBUFFERED=1
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2/tests_final.log; cat gpurun_out/r2/tests_final.log
python bench.py > gpurun_out/r2/final_c2.json 2> gpurun_out/r2/final_c2.err; python tools/bench_brief.py gpurun_out/r2/final_c2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/final_c2_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2/final_c2_reference.json
for c in c3 c5; do
  python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/final_$c.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r2/final_$c.json | head -2
done
python bench.py --cores 1000000 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/final_1Mcores.json 2>/dev/null; python tools/bench_brief.py gpurun_out/r2/final_1Mcores.json | head -2
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2/final_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2/final_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2/final_launches.csv > gpurun_out/r2/final_launches.txt; grep -v "at::" gpurun_out/r2/final_launches.txt | head -30

#!/usr/bin/env python3
"""Copies the evidence tools/round_profile.sh left in gpurun_out/ into profiles/ (tracked): bench lines, the
per-kernel launch list of one 50M-read step, and the ncu --set full metrics of the top kernels."""
import os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
for f in (f"{R}_bench_n1.json", f"{R}_bench_reference.json", f"{R}_scale_n2.json", f"{R}_scale_n4.json", f"{R}_scale_n8.json"):
    if os.path.exists(os.path.join(go, f)):
        shutil.copy(os.path.join(go, f), os.path.join(pr, f))
ll = os.path.join(go, f"{R}_launches.txt")
if os.path.exists(ll):
    with open(os.path.join(pr, f"{R}_launches_summary.txt"), "w") as o:
        o.write(f"# ncu launch list, {R} (cold-cache, serialised; compare SHARES). Command:\n"
                "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
                "python bench.py --reads 50000000 --steps 1 --warmup 1 --no-e2e --no-cpu\n"
                "# aggregated per kernel over the timed step (50M x 150 bp, 2048-core set) by tools/launch_summary.py\n")
        o.write(open(ll).read())
raw = os.path.join(go, f"{R}_top.raw.csv")
if os.path.exists(raw):
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), raw], stdout=subprocess.PIPE, text=True).stdout
    raw2 = os.path.join(go, f"{R}_sort.raw.csv")
    if os.path.exists(raw2):
        out += subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), raw2], stdout=subprocess.PIPE, text=True).stdout
    with open(os.path.join(pr, f"{R}_ncu_top_kernels.txt"), "w") as o:
        o.write(f"# ncu --set full --clock-control none --import-source on, one launch per kernel, {R}\n"
                "# workload for this capture: 5M x 150 bp (same code path as the 50M bench; ncu replays each kernel ~40x)\n\n")
        o.write(out)
print(sorted(os.listdir(pr)))

#!/usr/bin/env python3
"""Instruction / stall-sample share per block of SASS instructions from `ncu --page source --csv`."""
import csv, subprocess, sys
rep, step = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]; data = rows[2:]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("total inst", tot, "samples", tots)
acc = accs = 0
for i, r in enumerate(data):
    acc += int(r[ia]); accs += int(r[ismp])
    if (i + 1) % step == 0 or i == len(data) - 1:
        print(f"{max(i-step+1,0):5d}-{i:5d} inst {acc/tot*100:5.1f}% samples {accs/tots*100:5.1f}%  e.g. {data[max(i-step//2,0)][isrc].strip()[:60]}")
        acc = accs = 0

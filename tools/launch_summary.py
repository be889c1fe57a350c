#!/usr/bin/env python3
"""Aggregates an ncu launch list (csv, metrics gpu__time_duration.sum + dram bytes) per kernel over the LAST step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; data = rows[1:]
iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
L = collections.OrderedDict()
for r in data:
    d = L.setdefault(int(r[iid]), {"k": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im].startswith("gpu__time"):
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    else:
        d["rd" if "read" in r[im] else "wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
ids = list(L)
# the last step starts at the last scan kernel
starts = [i for i in ids if L[i]["k"].startswith("scan_smem_k") or L[i]["k"].startswith("scb::scan_smem_k")]
lo = starts[-1] if starts else ids[0]
agg = collections.OrderedDict()
for i in ids:
    if i < lo:
        continue
    d = L[i]; k = d["k"].split("(")[0].replace("scb::", "")[:58]
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("ms", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    gbs = (a[2] + a[3]) / (a[1] * 1e-3) / 1e9 if a[1] > 0 else 0
    print(f"{k:58s} n={a[0]:4d} {a[1]:8.3f} ms {a[1]/tot*100:5.1f}% rd={a[2]/1e9:6.2f} wr={a[3]/1e9:6.2f} GB {gbs:8.1f} GB/s")
print("total ms", round(tot, 3), "launches", sum(a[0] for a in agg.values()))

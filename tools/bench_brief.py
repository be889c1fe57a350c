#!/usr/bin/env python3
"""One-screen summary of a bench.py JSON line."""
import json
import sys

for p in sys.argv[1:]:
    try:
        d = json.loads([ln for ln in open(p).read().splitlines() if ln.startswith("{")][-1])
    except Exception as ex:  # noqa: BLE001
        print(p, "no JSON line:", ex)
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print(f"{p}: {d['value'] / 1e9:.3f} G reads/s, {d['ms_per_step']:.2f} ms/step, n_gpus {d['n_gpus']}, launches {d.get('gpu_launches')}, "
          f"path {d['pipeline_roofline']['frac_of_peak']:.3f} of HBM peak, workspace {d.get('workspace_bytes_per_flush', 0) / 2**30:.1f} GiB")
    print("   stages ms:", {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, "rounds", r.get("resolve_rounds"))
    if r.get("phase_wall_ms_rank0"):
        w = r["phase_wall_ms_rank0"]
        print("   wall ms (rank 0):", {k: round(v, 2) for k, v in w.items()}, "sum", round(sum(v for k, v in w.items() if not k.startswith("_")), 2))
    for br in (r.get("by_rank") or []):
        w = br.get("wall_ms") or {}
        print(f"   rank {br['rank']}: {br['ms_per_step']:.2f} ms/step; wall", {k: round(v, 1) for k, v in w.items()})
    if e:
        print("   e2e:", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in e.items() if k in ("value", "ms_per_step", "ms_per_step_min", "ms_submit_flush_copyout", "error", "skipped")},
              "serial", e.get("serial"), "piped", (e.get("pipelined") or {}).get("ms_per_step"))
    if c:
        print("   cpu:", c.get("value"), c.get("cores"), c.get("kind"), c.get("error"))
    if d.get("parity"):
        print("   parity:", d["parity"])
    print("   clocks:", d.get("clocks"))

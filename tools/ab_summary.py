#!/usr/bin/env python3
"""Table of the bench lines tools/round2_ab*.sh left in a directory: ms per step and per stage, one column per run."""
import glob
import json
import os
import sys


def last_json_line(path):
    try:
        for ln in reversed(open(path).read().splitlines()):
            ln = ln.strip()
            if ln.startswith("{"):
                return json.loads(ln)
    except Exception:
        pass
    return None


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ab"
    prefix = sys.argv[2] if len(sys.argv) > 2 else ""
    runs = {}
    for p in sorted(glob.glob(os.path.join(d, prefix + "*.json"))):
        name = os.path.basename(p)[:-5]
        if not prefix and name.startswith("n") and "_" in name and name[1:name.index("_")].isdigit():
            continue   # multi-GPU lines are listed with their own prefix
        j = last_json_line(p)
        if j is None:
            err = p[:-5] + ".err"
            tail = open(err).read()[-300:].replace("\n", " | ") if os.path.exists(err) else ""
            print(f"{name}: no bench line ({tail})")
            continue
        runs[name] = j
    if not runs:
        print("no bench lines in", d)
        return
    names = list(runs)
    stages = []
    for j in runs.values():
        for k in (j.get("roofline") or {}).get("stage_ms", {}):
            if k not in stages:
                stages.append(k)
    w = max(12, max(len(n) for n in names) + 1)
    print("".ljust(16) + "".join(n.rjust(w) for n in names))
    print("ms/step".ljust(16) + "".join(f"{runs[n]['ms_per_step']:.2f}".rjust(w) for n in names))
    print("G reads/s".ljust(16) + "".join(f"{runs[n]['value'] / 1e9:.3f}".rjust(w) for n in names))
    for s in stages:
        print(s.ljust(16) + "".join((f"{runs[n]['roofline']['stage_ms'].get(s, float('nan')):.2f}").rjust(w) for n in names))
    print("rounds".ljust(16) + "".join(str((runs[n].get("roofline") or {}).get("resolve_rounds")).rjust(w) for n in names))
    for n in names:
        e = runs[n].get("e2e")
        if e and e.get('value'):
            print(f"e2e {n}: {e['value'] / 1e6:.1f} M reads/s, {e['ms_per_step']:.0f} ms/step, submit/flush/copy-out {e.get('ms_submit_flush_copyout')}, "
                  f"serial {e.get('serial')}, pipelined {e.get('pipelined')}")


if __name__ == "__main__":
    main()

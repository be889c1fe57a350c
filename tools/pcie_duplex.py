#!/usr/bin/env python3
"""What the host link of this box gives the end-to-end arm: pinned H2D alone, D2H alone, both at once (bench.py's e2e bound)."""
import sys
import time

import torch

GB = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = GB << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if h2d:
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for name, a, b in (("warm", 1, 1), ("h2d only", 1, 0), ("d2h only", 0, 1), ("both", 1, 1), ("both", 1, 1)):
    t = run(a, b)
    print(f"{name:9s}: {t * 1e3:8.1f} ms  -> " + (f"h2d {GB * 1.0737 / t:6.1f} GB/s " if a else "") + (f"d2h {GB * 1.0737 / t:6.1f} GB/s" if b else ""))

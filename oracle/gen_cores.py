#!/usr/bin/env python3
"""Seeded synthetic core sets (the reference's patterns.bin is absent from the checkout,
/root/reference/.MISSING_LARGE_BLOBS).

Writes the two on-disk forms the reference loads:
  text   - whitespace separated cores, read by read_patterns_from_file (reads.cpp:379-410)
  binary - patterns.bin records {int16 len; int32 cnt; cnt x ceil(len/4) bytes}, each core the
           low bytes of a little-endian integer whose bits [2j+1:2j] hold base (len-1-j)
           (read_patterns, reads.cpp:330-377).
Core index (what ends up in .scalcer bucket headers) is the position in the file.

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import struct
import numpy as np

ALPHA = "ACGT"


def make_cores(seed, spec):
    """spec: list of (length, count). Returns list[str], grouped by ascending length, distinct."""
    rng = np.random.default_rng(seed)
    out = []
    for ln, cnt in spec:
        cnt = min(cnt, 4 ** ln)
        seen = set()
        while len(seen) < cnt:
            need = cnt - len(seen)
            codes = rng.integers(0, 4, size=(need * 2 + 8, ln), dtype=np.uint8)
            for row in codes:
                s = "".join(ALPHA[c] for c in row)
                if s not in seen:
                    seen.add(s)
                    out.append(s)
                    if len(seen) == cnt:
                        break
    return out


def parse_spec(s):
    # "8:512,9:256" -> [(8,512),(9,256)]
    return [(int(a), int(b)) for a, b in (t.split(":") for t in s.split(","))]


def write_text(path, cores):
    with open(path, "w") as f:
        for c in cores:
            f.write(c + "\n")


def write_binary(path, cores):
    by_len = {}
    for c in cores:
        by_len.setdefault(len(c), []).append(c)
    with open(path, "wb") as f:
        for ln in sorted(by_len):
            grp = by_len[ln]
            f.write(struct.pack("<hi", ln, len(grp)))
            sz = (ln + 3) // 4
            for c in grp:
                x = 0
                for ch in c:
                    x = (x << 2) | ALPHA.index(ch)
                f.write(x.to_bytes(8, "little")[:sz])


def binary_order(cores):
    """Core order as read_patterns() assigns indices for a binary file written by write_binary."""
    return sorted(cores, key=len) if False else [c for ln in sorted({len(c) for c in cores}) for c in cores if len(c) == ln]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--spec", default="8:256,9:128,10:128,11:64,12:64")
    ap.add_argument("--text")
    ap.add_argument("--binary")
    a = ap.parse_args()
    cores = make_cores(a.seed, parse_spec(a.spec))
    if a.text:
        write_text(a.text, cores)
    if a.binary:
        write_binary(a.binary, cores)
    print(len(cores))

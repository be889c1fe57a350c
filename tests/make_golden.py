#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference CLI (oracle/_ref/scalce, built
by oracle/Makefile from /root/reference) at -T 1 on seeded synthetic inputs.

The reference ships no tests or golden vectors, so these fixtures - outputs of the
reference itself - are what pins the oracle. Run in the build container (needs /root/reference):

    python tests/make_golden.py

Tiny cases store inputs and full output bytes; larger ones store the input's and outputs' SHA-256
(the input is regenerated from the seed and checked against its hash before use).
Every case is also round-tripped through the reference's own decompressor here.
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle.gen_cores import make_cores, write_text  # noqa: E402
from scalce_b200 import synth  # noqa: E402

CASES = [
    # name, n, L, kwargs
    dict(name="se60_tiny", n=400, L=60, seed=101, spec=[(6, 24), (7, 12), (8, 8)], store="full"),
    dict(name="pe60_40_tiny", n=300, L=60, L2=40, paired=True, seed=102, spec=[(6, 24), (8, 8)], store="full"),
    dict(name="nonames_tiny", n=300, L=50, seed=103, spec=[(6, 16), (7, 8)], use_names=False, store="full"),
    dict(name="se100_chunks", n=12000, L=100, seed=104, spec=[(8, 256), (9, 128), (10, 128), (11, 64), (12, 64)], bucket="1M", store="hash"),
    dict(name="pe100_chunks", n=9000, L=100, L2=75, paired=True, seed=105, spec=[(8, 256), (10, 128), (12, 64)], bucket="1M", store="hash"),
    dict(name="se300_long", n=3000, L=300, seed=106, spec=[(8, 128), (14, 64), (20, 32)], store="hash"),
    dict(name="se36_lower", n=20000, L=36, seed=107, spec=[(8, 256), (9, 128)], lower=0.05, store="hash"),
]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def case_inputs(c):
    cores = make_cores(c["seed"], c["spec"])
    b = synth.make_batch(c["n"], c["L"], seed=c["seed"], paired=c.get("paired", False), L2=c.get("L2"), lower_frac=c.get("lower", 0.0))
    synth.plant_cores(b, cores, seed=c["seed"] + 1, frac=0.5)
    return cores, b


def input_digest(cores, b):
    h = hashlib.sha256()
    h.update("\n".join(cores).encode())
    for a in (b.seq, b.qual, b.names, b.name_off, b.seq2, b.qual2):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    orc.build(ref=True)
    out_dir = os.path.join(ROOT, "tests", "golden")
    for c in CASES:
        cores, b = case_inputs(c)
        d = tempfile.mkdtemp(prefix="golden_")
        try:
            write_text(d + "/cores.txt", cores)
            paired = c.get("paired", False)
            synth.write_fastq(b, d + "/in_1.fastq", d + "/in_2.fastq" if paired else None)
            use_names = c.get("use_names", True)
            orc.run_reference_cli(d + "/in_1.fastq", d + "/ref", d + "/cores.txt", paired=paired, bucket=c.get("bucket", "4G"),
                                  no_names=None if use_names else "lib", tmpdir=d + "/tmp")
            files = {}
            for mate in range(1 + int(paired)):
                for ext in "nrq":
                    files[f"{mate + 1}{ext}"] = open(f"{d}/ref_{mate + 1}.scalce{ext}", "rb").read()
            # round trip through the reference's own decompressor (names only when kept)
            orc.run_reference_decompress(d + "/ref_1.scalcen", d + "/rt", d + "/cores.txt", paired=paired)
            rt = open(d + "/rt_1.fastq", "rb").read().split(b"\n")
            got = sorted(zip(rt[1::4], rt[3::4]))
            want = sorted((b.seq[i].tobytes().upper().replace(b"N", b"N"), b.qual[i].tobytes()) for i in range(b.n))
            # the decompressor restores N only where quality is 0 and upper-cases nothing: compare 2-bit classes
            def norm(s):
                return bytes(orc_code(ch) for ch in s)
            def orc_code(ch):
                return {67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}.get(ch, 0)
            ok = sorted((norm(s), q) for s, q in got) == sorted((norm(s), q) for s, q in want)
            meta = dict(name=c["name"], n=c["n"], L=c["L"], L2=c.get("L2", 0) or 0, paired=paired, seed=c["seed"], spec=c["spec"],
                        use_names=use_names, bucket=c.get("bucket", "4G"), lower=c.get("lower", 0.0), store=c["store"],
                        input_sha=input_digest(cores, b), roundtrip_ok=bool(ok), sizes={k: len(v) for k, v in files.items()},
                        sha={k: sha(v) for k, v in files.items()})
            arrays = dict(meta=np.frombuffer(repr(meta).encode(), dtype=np.uint8))
            if c["store"] == "full":
                arrays.update(seq=b.seq, qual=b.qual, names=b.names, name_off=b.name_off,
                              cores=np.frombuffer("\n".join(cores).encode(), dtype=np.uint8))
                if paired:
                    arrays.update(seq2=b.seq2, qual2=b.qual2)
                for k, v in files.items():
                    arrays["out_" + k] = np.frombuffer(v, dtype=np.uint8)
            np.savez_compressed(os.path.join(out_dir, c["name"] + ".npz"), **arrays)
            print(c["name"], "roundtrip", ok, {k: len(v) for k, v in files.items()})
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()

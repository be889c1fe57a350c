"""The drop-in, proven: oracle/_ref/scalce_scb is the reference CLI with ONLY the boosting transform replaced - a copy of
compress.cpp with the INTEGRATION.md glue applied (oracle/dropin/make_dropin.py: compress.cpp:673-715, 799-801, 732-733, 821,
834) linked with libscalce_b200.so and the reference's other objects, unmodified. For every golden fixture (outputs of the
UNMODIFIED CLI at -T 1, tests/make_golden.py) it must write byte-identical .scalce{n,r,q} files, and the unmodified CLI's
decompressor must round-trip them to the input."""
import ast
import glob
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.gen_cores import write_text
from scalce_b200 import synth
from tests.test_oracle_golden import _inputs, _load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
SCB_CLI = os.path.join(ROOT, "oracle", "_ref", "scalce_scb")


def _need_cli():
    if not os.path.exists(SCB_CLI):
        if os.path.isdir("/root/reference"):
            from oracle.dropin import make_dropin
            make_dropin.build()
        else:
            pytest.skip("oracle/_ref/scalce_scb was not built (needs /root/reference at build time)")


def _run(cli, fastq1, out_prefix, cores_txt, meta, tmpdir, extra=()):
    cmd = [cli, fastq1, "-T", "1", "-o", out_prefix, "-B", meta["bucket"], "-P", cores_txt, "-c", "no", "-A", "-t", tmpdir]
    if meta["paired"]:
        cmd += ["-r"]
    if not meta["use_names"]:
        cmd += ["-n", "lib"]
    r = subprocess.run(cmd + list(extra), stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=os.path.dirname(out_prefix))
    assert r.returncode == 0, r.stderr.decode(errors="replace")[-3000:]
    return r


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_patched_cli_writes_the_reference_files(path, tmp_path):
    _need_cli()
    z, meta = _load(path)
    cores, b = _inputs(z, meta)
    d = str(tmp_path)
    write_text(d + "/cores.txt", cores)
    synth.write_fastq(b, d + "/in_1.fastq", d + "/in_2.fastq" if meta["paired"] else None)
    _run(SCB_CLI, d + "/in_1.fastq", d + "/gpu", d + "/cores.txt", meta, d + "/tmp")
    for mate in range(1 + int(meta["paired"])):
        for ext in "nrq":
            k = f"{mate + 1}{ext}"
            data = open(f"{d}/gpu_{mate + 1}.scalce{ext}", "rb").read()
            assert len(data) == meta["sizes"][k], f"{k}: size {len(data)} != reference {meta['sizes'][k]}"
            assert hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{k}: bytes differ from the unmodified reference CLI's output"
    # round trip through the UNMODIFIED reference decompressor
    if os.path.exists(orc.REF_CLI):
        orc.run_reference_decompress(d + "/gpu_1.scalcen", d + "/rt", d + "/cores.txt", paired=meta["paired"])
        rt = open(d + "/rt_1.fastq", "rb").read().split(b"\n")
        code = {67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}
        norm = lambda s_: bytes(code.get(ch, 0) for ch in s_)      # the decompressor restores N only where the quality is 0: compare 2-bit classes
        got = sorted((norm(s_), q) for s_, q in zip(rt[1::4], rt[3::4]))
        want = sorted((norm(b.seq[i].tobytes()), b.qual[i].tobytes()) for i in range(b.n))
        assert got == want, "round trip through the reference decompressor does not reproduce the input reads"


def test_patched_cli_default_mode_equals_unmodified_cli(tmp_path):
    """Default mode (gzip + arithmetic coding of the qualities): the host stages downstream of the transform consume its
    streams unchanged, so the compressed files are identical too. Compared with a live run of the unmodified CLI."""
    _need_cli()
    if not os.path.exists(orc.REF_CLI):
        pytest.skip("unmodified reference CLI not built")
    from oracle.gen_cores import make_cores
    cores = make_cores(921, [(8, 256), (9, 128), (10, 64)])
    b = synth.make_batch(30000, 100, seed=921, paired=True, L2=80)
    synth.plant_cores(b, cores, seed=922, frac=0.5)
    d = str(tmp_path)
    write_text(d + "/cores.txt", cores)
    synth.write_fastq(b, d + "/in_1.fastq", d + "/in_2.fastq")
    for cli, pre in ((orc.REF_CLI, "ref"), (SCB_CLI, "gpu")):
        r = subprocess.run([cli, d + "/in_1.fastq", "-T", "1", "-o", d + "/" + pre, "-B", "2M", "-P", d + "/cores.txt", "-r", "-t", d + "/tmp_" + pre],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=d)
        assert r.returncode == 0, r.stderr.decode(errors="replace")[-3000:]
    for mate in (1, 2):
        for ext in "nrq":
            a = open(f"{d}/ref_{mate}.scalce{ext}", "rb").read()
            g = open(f"{d}/gpu_{mate}.scalce{ext}", "rb").read()
            assert a == g, f"mate {mate} .scalce{ext}: {len(a)} vs {len(g)} bytes"

"""The C++ host side of the boundary end to end on a B200 (scalce_b200/host/scb_boost.cpp: FASTQ -> SoA batches -> scb_submit /
scb_flush -> the reference's temp files / raw containers), against the oracle's chunk streams and the reference CLI fixtures.
(Ran green on a B200 as part of the opt-in suite in round 2, then promoted here.)"""
import os

import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("device_parse", [False, True], ids=["host_parse", "device_parse"])
@pytest.mark.parametrize("paired", [False, True])
def test_host_tool_temp_files_match_oracle(tmp_path, paired, device_parse):
    import subprocess
    from scalce_b200 import build as bld, synth
    tool = bld.build_host_tool()
    cores, b, q1, q2, _ = util.make_case(30000, 100, seed=181, paired=paired, L2=75 if paired else None)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, bucket_set_bytes=1 << 21)
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(b, f1, f2 if paired else None)
    (tmp_path / "cores.txt").write_text("\n".join(cores) + "\n")
    out = tmp_path / "out"
    out.mkdir()
    cmd = [tool, f1] + (["-r", f2] if paired else []) + ["-P", str(tmp_path / "cores.txt"), "-o", str(out), "-B", str(1 << 21), "--merged", "--batch", "7001"] + (["--device-parse"] if device_parse else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    nf = 6 if paired else 4
    assert o.n_chunks > 1
    for c in range(o.n_chunks):
        for k in range(nf):
            assert (out / f"t_{c:03d}_{k}.tmp").read_bytes() == o.stream(k, c), f"chunk {c} stream {k}"
    assert not (out / f"t_{o.n_chunks:03d}_0.tmp").exists()
    for k in range(nf):
        assert (out / f"merged_{k}.tmp").read_bytes() == o.stream(k, -1), f"merged stream {k}"


# ---- joint tie-break rounds inside one kernel per rank (SCB_SHARD_JOINT_KERNEL=1): needs one PROCESS per GPU ----------------


def test_host_tool_fastq_to_container_matches_reference_golden(tmp_path):
    """FASTQ -> scb_boost --container (C++ host stages + the CUDA transform + container assembly) == the files the unmodified
    reference CLI wrote (tests/golden), for every fixture."""
    import glob
    import hashlib
    import subprocess
    from scalce_b200 import build as bld, synth
    from tests import test_oracle_golden as tg
    tool = bld.build_host_tool()
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))):
        z, meta = tg._load(path)
        cores, b = tg._inputs(z, meta)
        d = tmp_path / os.path.basename(path)[:-4]
        d.mkdir()
        f1, f2 = str(d / "in_1.fastq"), str(d / "in_2.fastq")
        synth.write_fastq(b, f1, f2 if meta["paired"] else None)
        (d / "cores.txt").write_text("\n".join(cores) + "\n")
        bucket = {"4G": 4 << 30, "1M": 1 << 20}[meta["bucket"]]
        cmd = [tool, f1] + (["-r", f2] if meta["paired"] else []) + ["-P", str(d / "cores.txt"), "-B", str(bucket), "--container", str(d / "out"),
                                                                       "--library", "lib"] + ([] if meta["use_names"] else ["-n"])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        for mate in range(1 + int(meta["paired"])):
            for ext in "nrq":
                k = f"{mate + 1}{ext}"
                data = (d / f"out_{mate + 1}.scalce{ext}").read_bytes()
                assert hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{os.path.basename(path)} {k}: differs from the reference CLI output"

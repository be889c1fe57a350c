"""The CUDA path against the fixtures the UNMODIFIED reference CLI produced (tests/golden/*.npz): the merged streams
that come out of the C ABI are assembled into the .scalce{n,r,q} containers (host stage, out of scope: done here by
the oracle module's helper) and compared byte for byte / by hash with the reference's files."""
import glob
import hashlib
import os

import pytest

from oracle import oracle as orc
from tests.test_oracle_golden import GOLD, _inputs, _load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_reference_cli(path):
    from scalce_b200.binding import BoostTransform
    z, meta = _load(path)
    cores, b = _inputs(z, meta)
    off = orc.detect_phred_offset(b.qual)
    q1 = orc.quality_payload(b.qual, b.seq, off)
    q2 = orc.quality_payload(b.qual2, b.seq2, orc.detect_phred_offset(b.qual2)) if meta["paired"] else None
    bucket = {"4G": 4 << 30, "1M": 1 << 20}[meta["bucket"]]
    t = BoostTransform(cores, meta["L"], meta["L2"], use_names=meta["use_names"], paired=meta["paired"],
                       bucket_set_bytes=bucket, emit_merged=True)
    t.submit(b.seq, q1, b.names if meta["use_names"] else None, b.name_off if meta["use_names"] else None, b.seq2, q2)
    r = t.flush()
    if meta["bucket"] == "1M":
        assert r.n_chunks > 1
    s = [r.stream(k, -1) for k in range(6)]
    for mate in range(1 + int(meta["paired"])):
        fn, fr, fq = orc.assemble_container(s[3], s[0], s[1], s[2], cores, meta["L"], off, use_names=meta["use_names"], library=b"lib",
                                            paired=meta["paired"], mate=mate, reads2=s[4], quals2=s[5], L2=meta["L2"])
        for ext, data in (("n", fn), ("r", fr), ("q", fq)):
            k = f"{mate + 1}{ext}"
            assert len(data) == meta["sizes"][k], f"{k}: size {len(data)} != reference {meta['sizes'][k]}"
            assert hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{k}: bytes differ from the reference CLI output"
    t.close()

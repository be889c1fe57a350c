"""SURVEY.md 8(f2) on the device: scb_submit_fastq (FASTQ text -> pending batch: parse loop compress.cpp:614-671, output_name
names.cpp:48-62, output_quality qualities.cpp:177-204 incl. its input-order ac_freq3 / ac_freq4 statistics) against the oracle's
restatement, which tests/test_oracle_golden.py pins to the unmodified reference objects."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle.gen_cores import make_cores
from scalce_b200 import synth
from scalce_b200.binding import BoostTransform, ScbError
from tests import util

pytestmark = pytest.mark.gpu


def _fastq(b, tmp_path, tag):
    p1, p2 = str(tmp_path / f"{tag}_1.fastq"), str(tmp_path / f"{tag}_2.fastq")
    synth.write_fastq(b, p1, p2 if b.seq2 is not None else None)
    return open(p1, "rb").read(), (open(p2, "rb").read() if b.seq2 is not None else None)


@pytest.mark.parametrize("kw", [dict(n=20000, L=100, seed=601), dict(n=9000, L=36, seed=602, lower=0.05), dict(n=3000, L=300, seed=603, high_entropy=True),
                                dict(n=8000, L=100, seed=604, paired=True, L2=75), dict(n=1, L=50, seed=605), dict(n=33, L=17, seed=606)],
                         ids=lambda k: f"n{k['n']}_L{k['L']}")
def test_fastq_text_through_the_device_front_end(kw, tmp_path):
    """The flush fed by scb_submit_fastq (two texts one after the other, the second without its last newline) equals the oracle fed
    the same reads, and the quality statistics equal the oracle's."""
    kw = dict(kw)
    n, L = kw.pop("n"), kw.pop("L")
    paired = kw.get("paired", False)
    cores, b, q1, q2, off = util.make_case(n, L, **kw)
    cut = n // 3
    import copy
    parts = []
    for a, z in ((0, cut), (cut, n)):
        if z <= a:
            continue
        sub = synth.FastqBatch(b.seq[a:z], b.qual[a:z], b.names[b.name_off[a]:b.name_off[z]], b.name_off[a:z + 1] - b.name_off[a],
                               b.seq2[a:z] if paired else None, b.qual2[a:z] if paired else None)
        parts.append(_fastq(sub, tmp_path, f"p{a}"))
    L2 = b.seq2.shape[1] if paired else 0
    o = util.run_oracle(cores, b, q1, q2, paired=paired, bucket_set_bytes=1 << 20)
    t = BoostTransform(cores, L, L2, paired=paired, bucket_set_bytes=1 << 20)
    st = [orc.new_quality_stats(), orc.new_quality_stats()]
    total = 0
    for k, (t1, t2) in enumerate(parts):
        if k == len(parts) - 1:
            t1 = t1[:-1]                               # a last line without its newline
        total += t.submit_fastq(t1, t2, phred_offset=(off, off))
        orc.parse_fastq(t1, L, off, st[0])
        if paired:
            orc.parse_fastq(t2, L2, off, st[1])
    assert total == n
    r = t.flush()
    util.assert_same(o, t, r, paired=paired)
    for m in range(2 if paired else 1):
        f3, f4 = t.quality_stats(m)
        assert np.array_equal(f3.ravel(), st[m]["freq3"]), f"ac_freq3 of mate {m}"
        assert np.array_equal(f4.ravel(), st[m]["freq4"]), f"ac_freq4 of mate {m}"
    t.reset_counts()
    f3, f4 = t.quality_stats(0)
    assert int(f3.sum()) == 0 and int(f4.sum()) == 0


def test_names_cut_at_first_space_and_mixed_with_plain_submit(tmp_path):
    cores, b, q1, q2, off = util.make_case(5000, 80, seed=611)
    text = bytearray()
    nm = b.names.tobytes()
    for i in range(b.n):
        name = nm[b.name_off[i]:b.name_off[i + 1]]
        text += b"@" + name + (b" length=80 extra" if i % 3 == 0 else b"") + b"\n" + b.seq[i].tobytes() + b"\n+" + (name if i % 5 == 0 else b"") + b"\n" + b.qual[i].tobytes() + b"\n"
    o = util.run_oracle(cores, b, q1, q2)
    t = BoostTransform(cores, 80)
    half = 2000
    cut = bytes(text).split(b"\n")
    first = b"\n".join(cut[:4 * half]) + b"\n"
    assert t.submit_fastq(first) == half
    t.submit(b.seq[half:], q1[half:], b.names, b.name_off[half:])       # the rest through the plain batch entry point
    r = t.flush()
    util.assert_same(o, t, r)


def test_malformed_text_is_rejected():
    cores = make_cores(5, [(8, 32)])
    t = BoostTransform(cores, 20)
    ok = b"@r1\n" + b"A" * 20 + b"\n+\n" + b"I" * 20 + b"\n"
    assert t.submit_fastq(ok) == 1
    for bad in (ok + b"@r2\n" + b"A" * 20 + b"\n+\n",                       # not a whole record
                b"r1\n" + b"A" * 20 + b"\n+\n" + b"I" * 20 + b"\n",       # no '@'
                b"@r1\n" + b"A" * 19 + b"\n+\n" + b"I" * 20 + b"\n",      # short read
                b"@r1\n" + b"A" * 20 + b"\n+\n" + b"I" * 21 + b"\n",      # long quality line
                b"@" + b"x" * 300 + b"\n" + b"A" * 20 + b"\n+\n" + b"I" * 20 + b"\n",   # name > 255
                b"@r1\n" + b"A" * 20 + b"\n+\n" + b"\x7e" * 19 + b" \n"):   # quality below the offset
        with pytest.raises(ScbError):
            t.submit_fastq(bad)
    r = t.flush()
    assert r.n_reads == 1

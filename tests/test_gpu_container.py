"""SURVEY.md 8(f3) and (f4) on the device, against the oracle (whose two restatements are pinned to the unmodified reference
CLI in tests/test_oracle_golden.py): the .scalcer body assembled from merged meta + stream 1 (compress.cpp:345-384), and the
decompress-side inverse (decompress.cpp:331-352)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import util

pytestmark = pytest.mark.gpu

CASES = [dict(n=20000, L=100, seed=501), dict(n=12000, L=100, seed=502, bucket_set_bytes=1 << 20), dict(n=4000, L=300, seed=503),
         dict(n=9000, L=36, seed=504, lower=0.05), dict(n=6000, L=100, seed=505, paired=True, L2=75, bucket_set_bytes=1 << 20),
         dict(n=5000, L=17, seed=506), dict(n=3, L=50, seed=507)]


def _run(kw):
    kw = dict(kw)
    run_kw = {k: kw.pop(k) for k in list(kw) if k in ("bucket_set_bytes",)}
    n, L = kw.pop("n"), kw.pop("L")
    paired = kw.get("paired", False)
    cores, b, q1, q2, off = util.make_case(n, L, **kw)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, **run_kw)
    t, r = util.run_cuda(cores, b, q1, q2, paired=paired, **run_kw)
    return cores, b, q1, q2, off, o, t, r, L, paired


@pytest.mark.parametrize("kw", CASES, ids=[f"n{c['n']}_L{c['L']}" for c in CASES])
def test_reads_container_body(kw):
    cores, b, q1, q2, off, o, t, r, L, paired = _run(kw)
    fn, fr, fq = orc.assemble_container(o.stream(3), o.stream(0), o.stream(1), o.stream(2), cores, L, off, paired=paired,
                                        reads2=o.stream(4), quals2=o.stream(5), L2=b.seq2.shape[1] if paired else 0)
    body, sc, sr = t.assemble_reads(-1)
    assert body == fr[16:], f"body differs: {len(body)} vs {len(fr) - 16} bytes, first diff {util._first_diff(body, fr[16:])}"
    wc, wr = orc.segments_from_meta(o.stream(3), cores, L, paired)
    assert np.array_equal(sc, wc) and np.array_equal(sr, wr) and int(sr.sum()) == b.n
    # per flush chunk: the same assembly over that chunk's temp-file streams
    for c in range(o.n_chunks):
        body_c, sc_c, sr_c = t.assemble_reads(c)
        fr_c = orc.assemble_container(o.stream(3, c), o.stream(0, c), o.stream(1, c), o.stream(2, c), cores, L, off, paired=paired,
                                      reads2=o.stream(4, c), quals2=o.stream(5, c), L2=b.seq2.shape[1] if paired else 0)[1]
        assert body_c == fr_c[16:], f"chunk {c}"


@pytest.mark.parametrize("kw", CASES, ids=[f"n{c['n']}_L{c['L']}" for c in CASES])
def test_inverse_reads(kw):
    cores, b, q1, q2, off, o, t, r, L, paired = _run(kw)
    sc, sr = orc.segments_from_meta(o.stream(3), cores, L, paired)
    want_seq, want_q = o.inverse(o.stream(1), sc, sr, quals=o.stream(2), phred=off)
    got_seq, got_q = t.inverse_reads(r.stream(1), sc, sr, quals=r.stream(2), phred_offset=off)
    assert np.array_equal(got_seq, want_seq) and np.array_equal(got_q, want_q)
    if o.n_chunks > 1:
        return          # perm describes the chunk-major order; the merged streams used above are bucket-major (compress.cpp:68-198)
    # and the round trip itself: output row j is input read perm[j] with bases mapped to ACGT (N where the quality is 0)
    perm = r.debug()["perm"].astype(np.int64)
    code = np.zeros(256, dtype=np.uint8)
    for ch, v in ((b"C", 1), (b"c", 1), (b"G", 2), (b"g", 2), (b"T", 3), (b"t", 3)):
        code[ch[0]] = v
    exp = np.frombuffer(b"ACGT", dtype=np.uint8)[code[b.seq[perm]]]
    exp[q1[perm] == 0] = ord("N")
    assert np.array_equal(got_seq, exp)
    assert np.array_equal(got_q, b.qual[perm]) or np.array_equal(got_q, (q1[perm] + off).astype(np.uint8))
    # without qualities: no N restoration
    s2, q2o = t.inverse_reads(r.stream(1), sc, sr, quals=None)
    assert q2o is None and np.array_equal(s2, np.frombuffer(b"ACGT", dtype=np.uint8)[code[b.seq[perm]]])
    if paired:
        L2 = b.seq2.shape[1]
        m2, mq = t.inverse_reads(r.stream(4), sc, sr, quals=r.stream(5), mate=1, phred_offset=off)
        e2 = np.frombuffer(b"ACGT", dtype=np.uint8)[code[b.seq2[perm]]]
        e2[q2[perm] == 0] = ord("N")
        assert m2.shape == (b.n, L2) and np.array_equal(m2, e2)


def test_inverse_rejects_bad_segment_table():
    from scalce_b200.binding import ScbError
    cores, b, q1, q2, off, o, t, r, L, paired = _run(dict(n=500, L=60, seed=511))
    sc, sr = orc.segments_from_meta(o.stream(3), cores, L)
    bad = sc.copy(); bad[0] = len(cores) + 5
    with pytest.raises(ScbError):
        t.inverse_reads(r.stream(1), bad, sr)

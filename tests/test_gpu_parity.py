"""Parity of the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _case(n, L, **kw):
    run_kw = {k: kw.pop(k) for k in list(kw) if k in ("use_names", "use_quals", "bucket_set_bytes", "splits")}
    paired = kw.get("paired", False)
    cores, b, q1, q2, _ = util.make_case(n, L, **kw)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, **run_kw)
    t, r = util.run_cuda(cores, b, q1, q2, paired=paired, **run_kw)
    util.assert_same(o, t, r, paired=paired)
    return o, t, r


def test_single_end_100():
    _case(20000, 100, seed=1)


def test_multi_chunk():
    o, t, r = _case(20000, 100, seed=2, bucket_set_bytes=1 << 20)
    assert r.n_chunks > 2


def test_paired():
    _case(20000, 100, seed=3, paired=True)


def test_paired_multi_chunk_l2():
    _case(20000, 100, seed=4, paired=True, L2=75, bucket_set_bytes=1 << 20)


def test_no_names():
    _case(20000, 100, seed=5, use_names=False)


def test_no_quals():
    _case(5000, 100, seed=15, use_quals=False)


def test_long_reads_two_byte_end():
    _case(5000, 300, seed=6, bucket_set_bytes=1 << 20)


def test_reads_at_the_reference_line_cap():
    """2498 bases = the longest read the reference's line buffer holds (const.h:87); the oracle is pinned against the reference
    CLI at this length (test_oracle_golden.py). End markers up to 2498 need 12 bits in the per-read metadata words."""
    _case(600, 2498, seed=401, plant=0.9, bucket_set_bytes=1 << 20)
    _case(500, 2047, seed=402)
    _case(400, 2498, seed=403, paired=True, L2=2498, bucket_set_bytes=1 << 20)
    _case(700, 1000, seed=404, use_names=False)
    from scalce_b200.binding import BoostTransform, ScbError
    with pytest.raises(ScbError):
        BoostTransform(util.make_cores(1, util.DEFAULT_SPEC), 2499, 0)


def test_short_reads_lowercase():
    _case(20000, 36, seed=7, lower=0.05)


def test_long_cores():
    _case(20000, 100, seed=8, spec=[(8, 100), (14, 50), (20, 30), (32, 10)])


def test_many_n_and_duplicates():
    # heavy N content makes many identical keys -> exercises tie refinement on long runs
    cores, b, q1, q2, _ = util.make_case(4000, 100, seed=9, n_frac=0.3)
    b.seq[1000:1400] = b.seq[0]          # 400 identical reads
    b.seq[2000:2200, :60] = b.seq[1, :60]  # long common prefixes
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_all_identical_reads():
    cores, b, q1, q2, _ = util.make_case(3000, 64, seed=10)
    b.seq[:] = b.seq[0]
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_split_submits_and_second_flush_keeps_lifetime_counts():
    cores, b, q1, q2, _ = util.make_case(12000, 100, seed=11)
    o = util.run_oracle(cores, b, q1, q2, splits=[5000])
    t, r = util.run_cuda(cores, b, q1, q2, splits=[100, 5000])
    util.assert_same(o, t, r)
    for ci in (0, 5, 17):
        assert o.lifetime_count(ci) == t.lifetime_count(ci)


def test_tiny_inputs():
    for n in (1, 2, 31, 33):
        _case(n, 50, seed=12 + n)


def test_no_core_starting_with_some_base():
    # aho_output meets the root early when a base starts no core (reads.cpp:473-476)
    spec_cores = ["CCGTAGGT", "GGATTACA", "TTTTGGGA", "CGCGCGAT"]  # nothing starts with A
    from scalce_b200 import synth
    b = synth.make_batch(3000, 60, seed=21)
    synth.plant_cores(b, spec_cores, seed=22, frac=0.6)
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(spec_cores, b, q1, None)
    t, r = util.run_cuda(spec_cores, b, q1, None)
    util.assert_same(o, t, r)


def test_dense_core_set_many_candidates():
    # every 4-mer is a core -> every position hits, dozens of distinct candidates per read
    import itertools
    cores = ["".join(p) for p in itertools.product("ACGT", repeat=4)]
    from scalce_b200 import synth
    b = synth.make_batch(2000, 80, seed=23)
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, None)
    t, r = util.run_cuda(cores, b, q1, None)
    util.assert_same(o, t, r)


def test_million_reads_bucket_ids():
    cores, b, q1, q2, _ = util.make_case(300000, 100, seed=31, plant=0.0,
                                         spec=[(8, 1024), (9, 512), (10, 256), (11, 128), (12, 128)])
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_sequential_resolve_engine(monkeypatch):
    # the fallback engine for bucket counts too large for shared memory, forced on a small case
    monkeypatch.setenv("SCB_RESOLVE", "seq")
    o, t, r = _case(20000, 100, seed=41)
    assert t.resolve_rounds == 0


def test_dense_resolve_used_by_default():
    o, t, r = _case(50000, 100, seed=42)
    assert t.resolve_rounds > 0


def test_global_table_scan_engine(monkeypatch):
    # the scan path for automata too large for shared memory, forced on a small case
    monkeypatch.setenv("SCB_SCAN", "global")
    _case(20000, 100, seed=43)


def test_sparse_first_tie_round():
    # >= 65536 reads: the first tie-refinement round collects the (rare) tied positions sparsely
    cores, b, q1, q2, _ = util.make_case(70000, 100, seed=61, n_frac=0.01)
    b.seq[5000:5300] = b.seq[17]            # 300 identical reads
    b.seq[40000:40100, :70] = b.seq[3, :70]  # long common prefixes
    b.seq[69990:70000] = b.seq[69989]
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_mostly_tied_falls_back_to_dense_round():
    cores, b, q1, q2, _ = util.make_case(70000, 64, seed=62)
    b.seq[:] = b.seq[:7].repeat(10000, axis=0)   # 7 distinct reads
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_mid_size_tie_groups():
    # groups of 33..1024 tied reads (short suffixes after the core, duplicates) take the CTA-per-group path
    cores, b, q1, q2, _ = util.make_case(80000, 100, seed=63)
    for k, (lo, n_dup) in enumerate(((1000, 40), (3000, 300), (9000, 1000), (20000, 1500))):
        b.seq[lo:lo + n_dup] = b.seq[lo]
        b.seq[lo:lo + n_dup:3, 90:] = b.seq[lo + n_dup + 1, 90:]   # every third differs late in the read
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_reads_of_16_32_and_odd_lengths():
    # rows of one or two packed words (<= 32 bases): PW = 1 broke the reciprocal row index of the scan's pack phase in round 1
    # (ceil(2^32 / 1) does not fit 32 bits); odd row words and 2-byte end markers take the 4-byte staging path of the stream-1 writer
    _case(7000, 17, seed=193)
    _case(5000, 32, seed=197)
    _case(4000, 16, seed=198, spec=[(8, 200), (9, 100)])
    _case(9000, 40, seed=191, bucket_set_bytes=1 << 20)
    _case(4000, 272, seed=194, plant=0.9, bucket_set_bytes=1 << 20)
    _case(3000, 250, seed=196)


def test_long_and_ragged_names():
    # names of 0..60 bytes: every alignment of the stream-0 record in the writer's staging buffer
    cores, b, q1, q2, _ = util.make_case(12000, 100, seed=211)
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 61, size=b.n)
    lens[:50] = np.arange(50) % 34
    off = np.zeros(b.n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    b.names = rng.integers(33, 127, size=int(off[-1]), dtype=np.uint8)
    b.name_off = off
    o = util.run_oracle(cores, b, q1, q2, bucket_set_bytes=1 << 20)
    t, r = util.run_cuda(cores, b, q1, q2, bucket_set_bytes=1 << 20)
    util.assert_same(o, t, r)


def test_second_flush_with_large_populations_dense_engine():
    # two flushes on one handle: the second starts from large lifetime populations (g0 > 0 in the dense engine's extrapolation)
    cores, b, q1, q2, _ = util.make_case(600000, 100, seed=204, plant=0.0, spec=[(8, 1024), (9, 512), (10, 256), (11, 128), (12, 128)])
    o = util.run_oracle(cores, b, q1, q2, splits=[350000])
    t, r = util.run_cuda(cores, b, q1, q2, splits=[350000])
    util.assert_same(o, t, r)


@pytest.mark.parametrize("paired", [False, True])
def test_streaming_flush_equals_one_flush(paired):
    """scb_flush_closed: a job fed in pieces, with only the complete flush chunks emitted after every piece and the open chunk's
    reads kept pending, produces exactly the chunks of ONE flush over the whole input (= the reference, which carries total_size
    across reads, compress.cpp:702-713): same boundaries, same bytes, same lifetime counts."""
    from scalce_b200.binding import BoostTransform
    B = 1 << 20
    cores, b, q1, q2, _ = util.make_case(60000, 100, seed=701 + int(paired), paired=paired, L2=75 if paired else None)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, bucket_set_bytes=B)
    L2 = 75 if paired else 0
    t = BoostTransform(cores, 100, L2, paired=paired, bucket_set_bytes=B, emit_merged=False)
    nstream = 6 if paired else 4
    chunks, emitted = [], 0
    cuts = [0, 300, 7000, 7001, 30000, 30500, 60000]      # the first piece is smaller than a chunk: nothing closes
    for a, z in zip(cuts[:-1], cuts[1:]):
        t.submit(b.seq[a:z], q1[a:z], b.names, b.name_off[a:z + 1], b.seq2[a:z] if paired else None, q2[a:z] if paired else None)
        r = t.flush_closed()
        emitted += r.n_reads
        assert emitted <= z
        for c in range(r.n_chunks):
            chunks.append([r.stream(k, c) for k in range(nstream)])
    r = t.flush()
    emitted += r.n_reads
    for c in range(r.n_chunks):
        chunks.append([r.stream(k, c) for k in range(nstream)])
    assert emitted == b.n and len(chunks) == o.n_chunks and o.n_chunks > 8
    for c in range(o.n_chunks):
        for k in range(nstream):
            assert chunks[c][k] == o.stream(k, c), f"chunk {c} stream {k}"
    for ci in (0, 5, 17, 100):
        assert o.lifetime_count(ci) == t.lifetime_count(ci)
    assert o.unbucketed == t.unbucketed

"""Core sets of production size (the reference sizes patterns[] for 5-10 M cores, reads.cpp:336, 385): the automaton lives
in global memory / L2 (scan_big.cuh) and the tie-break runs on bucket-major candidate lists (resolve_sparse.cuh). Same bar
as everywhere: bit-exact against the oracle, per-read arrays and every stream byte. The engines are also forced onto the
small standard cases (SCB_TABLE=global, SCB_RESOLVE=sparse) so that every shape of the default suite crosses them."""
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


@pytest.fixture
def big_engines(monkeypatch):
    monkeypatch.setenv("SCB_TABLE", "global")
    monkeypatch.setenv("SCB_RESOLVE", "sparse")


def test_forced_engines_are_the_ones_running(big_engines):
    o, t, r = _case(20000, 100, seed=301)
    assert not t.table_info()["smem_resident"]
    assert t.resolve_engine == 1
    assert t.resolve_rounds >= 2


def test_forced_standard_shapes(big_engines):
    _case(20000, 100, seed=302, bucket_set_bytes=1 << 20)
    _case(12000, 100, seed=303, paired=True, L2=75, bucket_set_bytes=1 << 20)
    _case(20000, 36, seed=304, lower=0.05)
    _case(4000, 300, seed=305)
    _case(9000, 100, seed=306, use_names=False)
    _case(20000, 100, seed=307, spec=[(8, 100), (14, 50), (20, 30), (32, 10)])
    for n in (1, 2, 31, 33, 257):
        _case(n, 50, seed=310 + n)


def test_forced_short_reads(big_engines):
    _case(7000, 17, seed=193)
    _case(5000, 32, seed=197)
    _case(4000, 16, seed=198, spec=[(8, 200), (9, 100)])


def test_forced_lifetime_counts_across_flushes(big_engines):
    cores, b, q1, q2, _ = util.make_case(12000, 100, seed=311)
    o = util.run_oracle(cores, b, q1, q2, splits=[5000])
    t, r = util.run_cuda(cores, b, q1, q2, splits=[100, 5000])
    util.assert_same(o, t, r)
    for ci in (0, 5, 17):
        assert o.lifetime_count(ci) == t.lifetime_count(ci)
    # a second flush on the same handle starts from the populations of the first
    from oracle import oracle as orc
    cores2, b2, q12, _, _ = util.make_case(9000, 100, seed=312)
    o2 = orc.Oracle(cores, 100)
    o2.submit(b.seq, q1, b.names, b.name_off)
    o2.submit(b2.seq, q12, b2.names, b2.name_off)
    o2.finish()
    t.submit(b2.seq, q12, b2.names, b2.name_off)
    r2 = t.flush()
    d_o, d_c = o2.debug(), r2.debug()
    for k in ("node_id", "core", "end"):
        assert np.array_equal(d_o[k][12000:], d_c[k]), k


def test_forced_hot_buckets_span_many_tiles(big_engines):
    # few buckets, many reads: a bucket's pair list spans many 2048-pair tiles (carries of the segmented scan)
    cores, b, q1, q2, _ = util.make_case(120000, 64, seed=313, spec=[(6, 40)], plant=0.0)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_forced_dense_core_set_queue_overflow(big_engines):
    # every 4-mer is a core -> every position hits: the per-lane hit queue overflows (slow exact path, second attempt of
    # the candidate arrays)
    import itertools
    from scalce_b200 import synth
    cores = ["".join(p) for p in itertools.product("ACGT", repeat=4)]
    b = synth.make_batch(3000, 80, seed=314)
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, None)
    t, r = util.run_cuda(cores, b, q1, None)
    util.assert_same(o, t, r)


def test_forced_identical_reads(big_engines):
    cores, b, q1, q2, _ = util.make_case(3000, 64, seed=315)
    b.seq[:] = b.seq[0]
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    o = util.run_oracle(cores, b, q1, q2)
    t, r = util.run_cuda(cores, b, q1, q2)
    util.assert_same(o, t, r)


def test_forced_many_input_blocks(big_engines, monkeypatch):
    """The one-GPU sparse engine walks the input in blocks of reads (resolve_sparse.cuh); blocks of 1024 reads here, so that the
    standard shapes cross dozens of block boundaries, including blocks without a candidate and a ragged last block."""
    monkeypatch.setenv("SCB_SPARSE_BLOCK", "1024")
    o, t, r = _case(20000, 100, seed=351)
    assert t.resolve_engine == 1 and t.resolve_rounds > 19          # at least one round per block
    _case(12000, 100, seed=352, paired=True, L2=75, bucket_set_bytes=1 << 20)
    _case(20500, 36, seed=353, lower=0.05)
    _case(9000, 100, seed=354, spec=[(8, 100), (14, 50), (20, 30), (32, 10)])
    _case(3000, 24, seed=355, spec=[(12, 3)])                     # almost no read has a candidate: empty blocks
    for n in (1023, 1024, 1025, 2049):
        _case(n, 50, seed=360 + n)
    monkeypatch.setenv("SCB_SPARSE_BLOCK0", "256")                 # growing blocks: 256, 512, 1024, 2048, 2048, ...
    monkeypatch.setenv("SCB_SPARSE_BLOCK", "2048")
    _case(20000, 100, seed=357)
    cores, b, q1, q2, _ = util.make_case(12000, 100, seed=356)
    o = util.run_oracle(cores, b, q1, q2, splits=[5000])
    t, r = util.run_cuda(cores, b, q1, q2, splits=[100, 5000])
    util.assert_same(o, t, r)


def test_sharded_sparse_engine(big_engines):
    for world, kw in ((2, dict(n=30000, L=100, seed=321)), (4, dict(n=40000, L=100, seed=322, bucket_set_bytes=1 << 20)),
                      (3, dict(n=20000, L=100, seed=323, paired=True, L2=75, bounds=[0, 1000, 13000, 20000]))):
        run_kw = {k: kw.pop(k) for k in list(kw) if k in ("bucket_set_bytes", "bounds")}
        n, L = kw.pop("n"), kw.pop("L")
        paired = kw.get("paired", False)
        cores, b, q1, q2, _ = util.make_case(n, L, **kw)
        o = util.run_oracle(cores, b, q1, q2, paired=paired, **{k: v for k, v in run_kw.items() if k != "bounds"})
        ranks = util.run_sharded_loopback(cores, b, q1, q2, world, paired=paired, **run_kw)
        util.assert_sharded_same(o, ranks, paired=paired)


def _big_case(n, L, spec, seed):
    from scalce_b200 import synth
    cores = synth.make_core_set(spec, seed=seed)
    b = synth.make_batch(n, L, seed=seed + 1)
    q1 = util.orc.quality_payload(b.qual, b.seq, 33)
    return cores, b, q1


def test_million_cores_million_reads():
    """1 M cores of 10-14 bases (2.9 M automaton states) x 1 M reads x 100 bp, default engine choice."""
    cores, b, q1 = _big_case(1_000_000, 100, [(10, 100_000), (11, 200_000), (12, 400_000), (13, 200_000), (14, 100_000)], seed=331)
    o = util.run_oracle(cores, b, q1, None, bucket_set_bytes=64 << 20)
    t, r = util.run_cuda(cores, b, q1, None, bucket_set_bytes=64 << 20)
    info = t.table_info()
    assert info["n_buckets"] == 1_000_000 and not info["smem_resident"] and t.resolve_engine == 1
    util.assert_same(o, t, r)
    assert r.n_chunks > 2


def test_million_twelvemers_many_ties_sharded():
    """1 M cores of one length: every read has ~5 equal-level candidates, the tie-break decides everything; 3 ranks."""
    cores, b, q1 = _big_case(300_000, 100, [(12, 1_000_000)], seed=341)
    o = util.run_oracle(cores, b, q1, None)
    t, r = util.run_cuda(cores, b, q1, None)
    util.assert_same(o, t, r)
    ranks = util.run_sharded_loopback(cores, b, q1, None, 3)
    util.assert_sharded_same(o, ranks)

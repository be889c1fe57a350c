"""Parity against the oracle on the SHAPES of the BASELINE configs (bit-exact, whole output): C1 at its full size
(1M x 100 bp), and C3 / C4 / C5 scaled to what the C oracle finishes in seconds."""
import pytest

from tests import util

pytestmark = pytest.mark.gpu

SPEC = [(8, 1024), (9, 512), (10, 256), (11, 128), (12, 128)]   # the bench's 2048-core shape


def _case(n, L, **kw):
    run_kw = {k: kw.pop(k) for k in list(kw) if k in ("use_names", "use_quals", "bucket_set_bytes")}
    paired = kw.get("paired", False)
    cores, b, q1, q2, _ = util.make_case(n, L, spec=SPEC, plant=0.0, **kw)
    o = util.run_oracle(cores, b, q1, q2, paired=paired, **run_kw)
    t, r = util.run_cuda(cores, b, q1, q2, paired=paired, **run_kw)
    util.assert_same(o, t, r, paired=paired)
    return o, t, r


def test_c1_1m_x_100_single_end():
    o, t, r = _case(1_000_000, 100, seed=101, bucket_set_bytes=64 << 20)   # several flush chunks, as -B 64M
    assert r.n_chunks >= 3 and t.resolve_rounds > 0


def test_c3_shape_paired_150():
    _case(400_000, 150, seed=102, paired=True, bucket_set_bytes=64 << 20)


def test_c4_shape_short_reads_36():
    _case(1_500_000, 36, seed=103, bucket_set_bytes=48 << 20)


def test_c5_shape_wide_reads_250_high_entropy_quals():
    _case(400_000, 250, seed=104, high_entropy=True, bucket_set_bytes=96 << 20)


def test_c3_shape_sharded_four_ranks():
    cores, b, q1, q2, _ = util.make_case(200_000, 150, spec=SPEC, plant=0.0, seed=105, paired=True)
    o = util.run_oracle(cores, b, q1, q2, paired=True, bucket_set_bytes=32 << 20)
    ranks = util.run_sharded_loopback(cores, b, q1, q2, 4, paired=True, bucket_set_bytes=32 << 20)
    util.assert_sharded_same(o, ranks, paired=True)

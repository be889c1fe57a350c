"""Full-size properties (BASELINE configs[1]: 50M x 150 bp, the bench workload) that do not need the oracle to
finish 50M reads: permutation, payload gathers, emission order, determinism - plus the PREFIX property that ties the
full-size run to the oracle: the tie-break is sequential in input order, so the first m reads of the big run must get
exactly the buckets / end markers the oracle gives those m reads alone.
SCB_FULLSIZE_READS overrides the read count (default 50,000,000)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

N_FULL = int(os.environ.get("SCB_FULLSIZE_READS", "50000000"))
L = 150
ROOT = (1 << 30) - 1


@pytest.fixture(scope="module")
def big():
    import torch
    import bench
    from scalce_b200 import synth
    from scalce_b200.binding import BoostTransform
    cores = bench.headline_cores()
    d = synth.make_batch_cuda(N_FULL, L, seed=1, device="cuda:0")
    seq, qual, names, name_off = d["seq"], d["qual"], d["names"], d["name_off"]
    qual = torch.where(seq == ord("N"), torch.zeros_like(qual), qual - 33)
    tr = BoostTransform(cores, L, device=0, emit_merged=False)
    tr.submit_device(N_FULL, seq.data_ptr(), qual.data_ptr(), names.data_ptr(), name_off.data_ptr())
    res = tr.flush()
    yield dict(cores=cores, seq=seq, qual=qual, names=names, name_off=name_off, tr=tr, res=res)
    tr.close()


def test_perm_is_a_permutation(big):
    import torch
    perm = big["res"].torch_view("perm").to(torch.int64)
    assert perm.numel() == N_FULL
    seen = torch.zeros(N_FULL, dtype=torch.uint8, device=perm.device)
    seen.index_fill_(0, perm, 1)
    assert int(seen.sum(dtype=torch.int64)) == N_FULL
    assert int(perm.min()) == 0 and int(perm.max()) == N_FULL - 1


def test_quality_and_name_streams_are_row_gathers(big):
    import torch
    res = big["res"]
    perm = res.torch_view("perm").to(torch.int64)
    q_out = res.torch_view(2)
    assert q_out.numel() == N_FULL * L
    step = 5_000_000
    for s in range(0, N_FULL, step):       # blockwise: keeps the temporary at 750 MB
        e = min(N_FULL, s + step)
        want = big["qual"].index_select(0, perm[s:e])
        assert torch.equal(q_out[s * L:e * L].view(e - s, L), want), f"quality stream differs in output rows {s}..{e}"
    W = 13                                  # fixed-width names SYN.%09d -> records [13][13 bytes]
    n_out = res.torch_view(0)
    assert n_out.numel() == N_FULL * (W + 1)
    rec = n_out.view(N_FULL, W + 1)
    assert bool((rec[:, 0] == W).all())
    nm = big["names"].view(N_FULL, W)
    for s in range(0, N_FULL, step):
        e = min(N_FULL, s + step)
        assert torch.equal(rec[s:e, 1:], nm.index_select(0, perm[s:e]))


def test_emission_order(big):
    """chunk-major, then bucket node id ascending with the no-core bucket last (reads.cpp:466-499), then the
    suffix key, then input order (stable radix sort, reads.cpp:547-634) - over the whole output, 2M-read windows."""
    import torch
    res = big["res"]
    perm = res.torch_view("perm").to(torch.int64)
    node = res.torch_view("node_id").to(torch.int64)
    chunk = res.torch_view("chunk").to(torch.int64)
    end = res.torch_view("end").to(torch.int64)
    b = node[perm]
    c = chunk[perm]
    seg = c * (1 << 31) + b                 # root id 2^30-1 sorts last among node ids
    assert bool((seg[1:] >= seg[:-1]).all()), "segments (chunk, bucket) are not in emission order"
    assert res.n_chunks == int(chunk.max()) + 1
    # inside segments: key = s[end..L) padded with A, N -> A; ties by input index
    m = min(2_000_000, N_FULL)
    for start in range(0, N_FULL, m - 1):              # every output position: windows overlap by one read
        if start + 1 >= N_FULL:
            break
        mm = min(m, N_FULL - start)
        idx = perm[start:start + mm]
        rows = big["seq"].index_select(0, idx)
        code = torch.zeros_like(rows)
        lo = rows | 0x20
        code[lo == ord("c")] = 1
        code[lo == ord("g")] = 2
        code[lo == ord("t")] = 3
        e = end[idx]
        pos = e[:, None] + torch.arange(L, device=rows.device)[None, :]
        key = torch.where(pos < L, torch.gather(code, 1, pos.clamp(max=L - 1)), torch.zeros_like(code))
        same_seg = seg[start + 1:start + mm] == seg[start:start + mm - 1]
        a, z = key[:-1], key[1:]
        diff = a != z
        has = diff.any(dim=1)
        first = torch.argmax(diff.to(torch.uint8), dim=1)
        av = torch.gather(a, 1, first[:, None])[:, 0]
        zv = torch.gather(z, 1, first[:, None])[:, 0]
        ok = torch.where(has, av < zv, idx[:-1] < idx[1:])
        assert bool((ok | ~same_seg).all()), f"in-bucket order violated in window at {start}"
        del rows, code, lo, pos, key, a, z, diff


def test_bucket_histogram_matches_lifetime_counts(big):
    import torch
    res, tr = big["res"], big["tr"]
    core = res.torch_view("core").to(torch.int64)
    hist = torch.bincount(core[core >= 0], minlength=len(big["cores"])).cpu().numpy()
    for ci in (0, 1, 77, 1024, 2047):
        assert tr.lifetime_count(ci) == int(hist[ci])
    assert tr.lifetime_count(-1) == int((core < 0).sum())
    assert tr.unbucketed == int((core < 0).sum())


def test_stream1_size_and_meta_totals(big):
    import torch
    res = big["res"]
    core = res.torch_view("core").to(torch.int64)
    lens = torch.tensor([len(c) for c in big["cores"]] + [0], device=core.device)
    lvl = lens[torch.where(core >= 0, core, torch.full_like(core, len(big["cores"])))]
    want = int(((L - lvl + 3) // 4 + 1).sum())
    assert res.chunk_off[1][-1] == want
    meta = np.frombuffer(b"".join(res.stream(3, c) for c in range(res.n_chunks)), dtype=np.uint8).reshape(-1, 32)
    tR = meta[:, 16:24].copy().view(np.int64).sum()
    tQ = meta[:, 24:32].copy().view(np.int64).sum()
    assert int(tR) == want and int(tQ) == N_FULL * L


def test_stream1_content_windows(big):
    """Stream 1 byte for byte on three 500k-read windows of the output: every record must be the rotated 2-bit read + end marker
    that reads.cpp:432-461 / 128-130 define for (read, core level, end) - tests/util.py::stream1_window_matches, itself checked
    against the oracle's stream on the CPU (tests/test_host_cpu.py)."""
    import torch
    from tests import util
    res = big["res"]
    perm = res.torch_view("perm").to(torch.int64)
    core = res.torch_view("core").to(torch.int64)
    lens = torch.tensor([len(c) for c in big["cores"]] + [0], device=core.device)
    lv = lens[torch.where(core >= 0, core, torch.full_like(core, len(big["cores"])))]
    end = res.torch_view("end").to(torch.int64)
    s1 = res.torch_view(1)
    m = min(500_000, N_FULL)
    for start in (0, max(0, N_FULL // 2 - m // 2), N_FULL - m):
        assert util.stream1_window_matches(s1, big["seq"], perm, lv, end, L, start, m), f"stream 1 differs in the window at output position {start}"


def test_prefix_of_the_big_run_equals_oracle_on_the_prefix(big):
    """Sequential semantics: reads 0..m-1 are decided before anything later exists, so their bucket ids and end
    markers in the 50M-read run equal the oracle's on those m reads alone (flush chunks too: same byte budget)."""
    from oracle import oracle as orc
    m = min(1_000_000, N_FULL)
    seq = big["seq"][:m].cpu().numpy()
    qual = big["qual"][:m].cpu().numpy()
    names = big["names"][:m * 13].cpu().numpy()
    off = big["name_off"][:m + 1].cpu().numpy()
    o = orc.Oracle(big["cores"], L)
    o.submit(seq, qual, names, off)
    o.finish()
    d = o.debug()
    res = big["res"]
    for k in ("node_id", "core", "end"):
        got = res.torch_view(k)[:m].cpu().numpy()
        bad = np.nonzero(d[k] != got)[0]
        assert bad.size == 0, f"{k} differs from the oracle at reads {bad[:5]}"


def test_every_read_of_the_big_run_equals_the_oracle(big):
    """ALL reads of the full-size run (not a prefix): bucket id, core, end marker and flush chunk against the oracle's
    streaming form (orc_assign: aho_search + size accounting + bin_size++, reads.cpp:413-429, compress.cpp:673-715; the same
    code path as the full oracle, tests/test_host_cpu.py) - this covers the late geometric blocks of the tie-break (12.5 M
    and 25 M reads), their replay rounds and the u32 population counters. ~80 s of one host core at 50 M reads."""
    from oracle import oracle as orc
    res = big["res"]
    o = orc.Oracle(big["cores"], L)
    step = 2_500_000
    nbad = 0
    for s in range(0, N_FULL, step):
        e = min(N_FULL, s + step)
        seq = big["seq"][s:e].cpu().numpy()
        off = big["name_off"][s:e + 1].cpu().numpy()
        d = o.assign(seq, off)
        for k in ("node_id", "core", "end", "chunk"):
            got = res.torch_view(k)[s:e].cpu().numpy()
            bad = np.nonzero(d[k] != got)[0]
            assert bad.size == 0, f"{k} differs from the oracle at reads {bad[:5] + s}: oracle {d[k][bad[:5]]} cuda {got[bad[:5]]}"
    assert o.unbucketed == big["tr"].unbucketed


def test_rerun_is_deterministic(big):
    import torch
    tr, res = big["tr"], big["res"]
    sums = [int(res.torch_view(k).sum(dtype=torch.int64)) for k in (0, 1, 2, 3)]
    p0 = res.torch_view("perm").clone()
    tr.reset_counts()
    tr.submit_device(N_FULL, big["seq"].data_ptr(), big["qual"].data_ptr(), big["names"].data_ptr(), big["name_off"].data_ptr())
    r2 = tr.flush()
    assert [int(r2.torch_view(k).sum(dtype=torch.int64)) for k in (0, 1, 2, 3)] == sums
    assert torch.equal(r2.torch_view("perm"), p0)

"""Shared helpers for the parity tests: one seeded case -> oracle result and CUDA result."""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc
from oracle.gen_cores import make_cores
from scalce_b200 import synth

DEFAULT_SPEC = [(8, 256), (9, 128), (10, 128), (11, 64), (12, 64)]


def make_case(n, L, spec=DEFAULT_SPEC, seed=1, paired=False, L2=None, plant=0.5, lower=0.0, n_frac=0.001, high_entropy=False):
    cores = make_cores(seed, spec)
    b = synth.make_batch(n, L, seed=seed, paired=paired, L2=L2, lower_frac=lower, n_frac=n_frac, high_entropy=high_entropy)
    if plant:
        synth.plant_cores(b, cores, seed=seed + 1, frac=plant)
    off = orc.detect_phred_offset(b.qual)
    q1 = orc.quality_payload(b.qual, b.seq, off)
    q2 = orc.quality_payload(b.qual2, b.seq2, orc.detect_phred_offset(b.qual2)) if paired else None
    return cores, b, q1, q2, off


def run_oracle(cores, b, q1, q2, *, use_names=True, paired=False, use_quals=True, bucket_set_bytes=4 << 30, splits=None):
    L1 = b.seq.shape[1]
    L2 = b.seq2.shape[1] if paired else 0
    o = orc.Oracle(cores, L1, L2, use_names=use_names, paired=paired, use_quals=use_quals, bucket_set_bytes=bucket_set_bytes)
    _submit_split(o.submit, b, q1, q2, paired, splits)
    o.finish()
    return o


def _submit_split(submit, b, q1, q2, paired, splits):
    n = b.seq.shape[0]
    cuts = [0] + list(splits or []) + [n]
    for a, z in zip(cuts[:-1], cuts[1:]):
        if z <= a:
            continue
        no = b.name_off[a:z + 1]
        submit(b.seq[a:z], q1[a:z] if q1 is not None else None, b.names, no,
               b.seq2[a:z] if paired else None, q2[a:z] if (paired and q2 is not None) else None)


def run_cuda(cores, b, q1, q2, *, use_names=True, paired=False, use_quals=True, bucket_set_bytes=4 << 30, splits=None,
             emit_merged=True):
    from scalce_b200.binding import BoostTransform
    L1 = b.seq.shape[1]
    L2 = b.seq2.shape[1] if paired else 0
    t = BoostTransform(cores, L1, L2, use_names=use_names, paired=paired, use_quals=use_quals,
                       bucket_set_bytes=bucket_set_bytes, emit_merged=emit_merged)
    _submit_split(t.submit, b, q1 if use_quals else None, q2 if use_quals else None, paired, splits)
    r = t.flush()
    return t, r


def assert_same(o, t, r, paired=False, check_merged=True):
    dbg_o, dbg_c = o.debug(), r.debug()
    for k in ("node_id", "core", "end", "chunk"):
        bad = np.nonzero(dbg_o[k] != dbg_c[k])[0]
        assert bad.size == 0, f"per-read {k} differs at {bad[:5]}: oracle {dbg_o[k][bad[:5]]} cuda {dbg_c[k][bad[:5]]}"
    assert o.n_chunks == r.n_chunks
    streams = [0, 1, 2, 3] + ([4, 5] if paired else [])
    for c in range(o.n_chunks):
        for k in streams:
            a, b = o.stream(k, c), r.stream(k, c)
            assert a == b, f"chunk {c} stream {k}: oracle {len(a)} B, cuda {len(b)} B, first diff {_first_diff(a, b)}"
    if check_merged:
        for k in streams:
            a, b = o.stream(k, -1), r.stream(k, -1)
            assert a == b, f"merged stream {k}: oracle {len(a)} B, cuda {len(b)} B, first diff {_first_diff(a, b)}"
    assert o.unbucketed == t.unbucketed


def _first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], dtype=np.uint8) != np.frombuffer(b[:n], dtype=np.uint8)
    nz = np.nonzero(x)[0]
    return int(nz[0]) if nz.size else n


def run_sharded_loopback(cores, b, q1, q2, world, *, use_names=True, paired=False, use_quals=True, bucket_set_bytes=4 << 30,
                         emit_merged=True, bounds=None, device=0, cpp=False):
    """The sharded path with `world` ranks as threads of this process on ONE GPU (LoopbackComm).
    Returns [(transform, sharded, result)] per rank."""
    import threading
    from scalce_b200.binding import BoostTransform
    from scalce_b200.shard import CShardedTransform, LoopbackComm, ShardedTransform, shard_bounds
    n = b.seq.shape[0]
    L1 = b.seq.shape[1]
    L2 = b.seq2.shape[1] if paired else 0
    bounds = bounds or shard_bounds(n, world)
    comms = LoopbackComm.make(world, device)
    out = [None] * world
    errs = []

    def worker(r):
        try:
            t = BoostTransform(cores, L1, L2, use_names=use_names, paired=paired, use_quals=use_quals,
                               bucket_set_bytes=bucket_set_bytes, emit_merged=emit_merged, device=device)
            a, z = bounds[r], bounds[r + 1]
            if z > a:
                t.submit(b.seq[a:z], q1[a:z] if (use_quals and q1 is not None) else None, b.names, b.name_off[a:z + 1],
                         b.seq2[a:z] if paired else None, q2[a:z] if (paired and use_quals and q2 is not None) else None)
            st = CShardedTransform(t, comms[r]) if cpp else ShardedTransform(t, comms[r])   # cpp: the sequence runs inside scb_shard_flush
            out[r] = (t, st, st.flush())
        except BaseException as e:   # noqa: BLE001 - unblock the other ranks, then re-raise in the caller
            errs.append(e)
            comms[r].s.barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    if errs:
        real = [e for e in errs if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or errs)[0]
    return out


def assert_sharded_same(o, ranks, paired=False, check_merged=True):
    """Rank-order concatenation of the ranks' outputs == the oracle over the whole input."""
    dbg_o = o.debug()
    dbg = [res.debug(res.n_local) for _, _, res in ranks]
    for k in ("node_id", "core", "end", "chunk"):
        got = np.concatenate([d[k] for d in dbg])
        bad = np.nonzero(dbg_o[k] != got)[0]
        assert bad.size == 0, f"per-read {k} differs at {bad[:5]}: oracle {dbg_o[k][bad[:5]]} sharded {got[bad[:5]]}"
    for _, _, res in ranks:
        assert res.n_chunks == o.n_chunks, (res.n_chunks, o.n_chunks)
    streams = [0, 1, 2, 3] + ([4, 5] if paired else [])
    for c in range(o.n_chunks):
        for k in streams:
            a = o.stream(k, c)
            g = b"".join(res.stream(k, c) for _, _, res in ranks)
            assert a == g, f"chunk {c} stream {k}: oracle {len(a)} B, sharded {len(g)} B, first diff {_first_diff(a, g)}"
    if check_merged:
        for k in streams:
            a = o.stream(k, -1)
            g = b"".join(res.stream(k, -1) for _, _, res in ranks)
            assert a == g, f"merged stream {k}: oracle {len(a)} B, sharded {len(g)} B, first diff {_first_diff(a, g)}"
    assert o.unbucketed == sum(t.unbucketed for t, _, _ in ranks)


def expected_stream1_records(seq_rows, lv, end, L):
    """torch (CPU or CUDA): the stream-1 records of the given reads in the given order - output_read(read, dest, end - level, level)
    (reads.cpp:432-461: bases [end, L) then [0, end - level), 4 per byte MSB first, zero padded) followed by the 1-byte end marker
    (L <= 255, reads.cpp:128-130). seq_rows uint8 [m, L] ASCII; lv, end int64 [m]. Returns (records uint8 [m, ceil(L/4) + 1] zero
    padded, sizes int64 [m])."""
    import torch
    assert L <= 255
    m, dev = seq_rows.shape[0], seq_rows.device
    lo = seq_rows | 0x20
    code = torch.zeros_like(seq_rows)
    code[lo == ord("c")] = 1
    code[lo == ord("g")] = 2
    code[lo == ord("t")] = 3
    j = torch.arange(L, device=dev)[None, :]
    tail, total = (L - end)[:, None], (L - lv)[:, None]
    src = torch.where(j < tail, end[:, None] + j, j - tail)
    out_code = torch.where(j < total, torch.gather(code, 1, src.clamp(0, L - 1)), torch.zeros_like(code))
    nbmax = (L + 3) // 4
    q = torch.nn.functional.pad(out_code, (0, nbmax * 4 - L)).view(m, nbmax, 4).to(torch.int32)
    by = ((q[:, :, 0] << 6) | (q[:, :, 1] << 4) | (q[:, :, 2] << 2) | q[:, :, 3]).to(torch.uint8)
    nbytes = (L - lv + 3) // 4
    rec = torch.zeros((m, nbmax + 1), dtype=torch.uint8, device=dev)
    rec[:, :nbmax] = by
    rec.scatter_(1, nbytes[:, None], (end & 0xff).to(torch.uint8)[:, None])
    return rec, nbytes + 1


def stream1_window_matches(stream1, seq, perm, lv_in, end_in, L, start, m):
    """True if the records of output positions [start, start + m) in `stream1` (uint8 tensor, all chunks in output order) are what
    expected_stream1_records says. perm int64 [n] output position -> input read; lv_in / end_in int64 [n] per INPUT read."""
    import torch
    sizes = (L - lv_in[perm] + 3) // 4 + 1
    off = torch.cumsum(sizes, 0) - sizes
    idx = perm[start:start + m]
    want, sz = expected_stream1_records(seq.index_select(0, idx), lv_in[idx], end_in[idx], L)
    col = torch.arange(want.shape[1], device=want.device)[None, :]
    pos = (off[start:start + m, None] + col).clamp(max=stream1.numel() - 1)
    got = torch.where(col < sz[:, None], stream1[pos], torch.zeros_like(want))
    return bool(torch.equal(got, want)) and int(off[-1] + sizes[-1]) == stream1.numel()

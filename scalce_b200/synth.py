"""Seeded synthetic fixed-length FASTQ batches.

Bases i.i.d. uniform over ACGT with a small fraction of 'N' (quality forced to the offset
character there, as real instruments do); names ``SYN.<i>``; qualities phred+33, either a
peaked low-entropy distribution (default) or i.i.d. uniform 2..41 ("high entropy", config C5).

Two producers with the same schema:
  * :func:`make_batch`      numpy, host memory  - tests, golden fixtures, FASTQ files
  * :func:`make_batch_cuda` torch, device memory - bench-scale inputs generated in HBM

A batch is the structure-of-arrays the host side of the boundary builds from FASTQ
(the role of the parse loop, compress.cpp:614-671 in the reference):
  seq   uint8 [n, L]   ASCII bases, row pitch L (no newline)
  qual  uint8 [n, L]   ASCII qualities
  names bytes          concatenated names (without '@'), name_off int64 [n+1]
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class FastqBatch:
    seq: np.ndarray        # uint8 [n, L]
    qual: np.ndarray       # uint8 [n, L]
    names: np.ndarray      # uint8 [total]
    name_off: np.ndarray   # int64 [n+1]
    seq2: np.ndarray | None = None
    qual2: np.ndarray | None = None

    @property
    def n(self):
        return self.seq.shape[0]


def _quals(rng, n, L, high_entropy):
    if high_entropy:
        q = rng.integers(2, 42, size=(n, L), dtype=np.uint8)
    else:
        # peaked: mostly 36..40, occasional low values; a slow downward drift along the read
        base = rng.choice(np.array([40, 40, 40, 39, 38, 37, 36, 30, 20, 2], dtype=np.uint8), size=(n, L))
        drop = (rng.random((n, 1)) * np.linspace(0, 6, L)[None, :]).astype(np.uint8)
        q = np.maximum(base.astype(np.int16) - drop, 2).astype(np.uint8)
    return (q + 33).astype(np.uint8)


def _mate(rng, n, L, n_frac, high_entropy, lower_frac=0.0):
    seq = _BASES[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    qual = _quals(rng, n, L, high_entropy)
    if n_frac > 0:
        m = rng.random((n, L)) < n_frac
        seq[m] = ord("N")
        qual[m] = 33
    if lower_frac > 0:
        m = rng.random((n, L)) < lower_frac
        seq[m] |= 0x20  # lower case, accepted by getval (const.cpp:47-49)
    return seq, qual


def make_names(n, start=0, prefix="SYN."):
    idx = np.arange(start, start + n)
    strs = np.char.add(prefix, idx.astype(str)).astype("S")
    lens = np.char.str_len(strs).astype(np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    w = strs.dtype.itemsize
    flat = np.frombuffer(strs.tobytes(), dtype=np.uint8).reshape(n, w)
    mask = np.arange(w)[None, :] < lens[:, None]
    return flat[mask].copy(), off


def make_batch(n, L, seed=1, paired=False, L2=None, n_frac=0.001, high_entropy=False,
               lower_frac=0.0, name_start=0) -> FastqBatch:
    rng = np.random.default_rng(seed)
    seq, qual = _mate(rng, n, L, n_frac, high_entropy, lower_frac)
    names, off = make_names(n, name_start)
    b = FastqBatch(seq, qual, names, off)
    if paired:
        b.seq2, b.qual2 = _mate(rng, n, L2 or L, n_frac, high_entropy, lower_frac)
    return b


def plant_cores(batch: FastqBatch, cores, seed=3, frac=0.5):
    """Overwrite a random window of a fraction of reads with a random core so that long cores
    (which i.i.d. reads almost never contain) get exercised."""
    rng = np.random.default_rng(seed)
    n, L = batch.seq.shape
    pick = np.nonzero(rng.random(n) < frac)[0]
    which = rng.integers(0, len(cores), size=pick.size)
    for r, w in zip(pick, which):
        c = np.frombuffer(cores[w].encode(), dtype=np.uint8)
        if c.size > L:
            continue
        p = rng.integers(0, L - c.size + 1)
        batch.seq[r, p:p + c.size] = c
        # a planted base over a former 'N' must not keep quality 0 (the decompressor restores 'N' there)
        q = batch.qual[r, p:p + c.size]
        q[q == 33] = 35
    return batch


def write_fastq(batch: FastqBatch, path1, path2=None):
    """Materialise the batch as FASTQ (names get /1 /2 suffix-free form; '+' line bare)."""
    def dump(path, seq, qual):
        n, L = seq.shape
        with open(path, "wb") as f:
            nm = batch.names.tobytes()
            off = batch.name_off
            # assemble in blocks to keep python overhead low
            blk = 65536
            for s in range(0, n, blk):
                e = min(n, s + blk)
                parts = []
                for i in range(s, e):
                    parts.append(b"@" + nm[off[i]:off[i + 1]] + b"\n" + seq[i].tobytes() + b"\n+\n" + qual[i].tobytes() + b"\n")
                f.write(b"".join(parts))
    dump(path1, batch.seq, batch.qual)
    if path2 is not None and batch.seq2 is not None:
        dump(path2, batch.seq2, batch.qual2)


def make_batch_cuda(n, L, seed=1, paired=False, L2=None, n_frac=0.001, high_entropy=False, device="cuda"):
    """Same schema generated directly in HBM with torch (bench-scale inputs; never on the timed path).

    Returns dict of torch tensors: seq,qual [n,L] uint8; names uint8 [total]; name_off int64 [n+1];
    optionally seq2, qual2. Names are fixed-width zero-padded ``SYN.%09d`` so offsets are closed-form.
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)

    def mate(LL):
        seq = torch.empty((n, LL), dtype=torch.uint8, device=device)
        qual = torch.empty((n, LL), dtype=torch.uint8, device=device)
        step = max(1, (1 << 28) // LL)
        qtab = torch.tensor([40, 40, 40, 39, 38, 37, 36, 30, 20, 2], dtype=torch.uint8, device=device)
        for s in range(0, n, step):
            e = min(n, s + step)
            r = torch.randint(0, 4, (e - s, LL), generator=g, device=device, dtype=torch.uint8)
            sq = lut[r.long()]
            if high_entropy:
                q = torch.randint(2, 42, (e - s, LL), generator=g, device=device, dtype=torch.uint8)
            else:
                q = qtab[torch.randint(0, 10, (e - s, LL), generator=g, device=device).long()]
            q = q + 33
            if n_frac > 0:
                m = torch.rand((e - s, LL), generator=g, device=device) < n_frac
                sq[m] = ord("N")
                q[m] = 33
            seq[s:e] = sq
            qual[s:e] = q
            del r, sq, q
        return seq, qual

    out = {}
    out["seq"], out["qual"] = mate(L)
    if paired:
        out["seq2"], out["qual2"] = mate(L2 or L)
    W = 13  # "SYN." + 9 digits
    idx = torch.arange(n, device=device, dtype=torch.int64)
    names = torch.empty((n, W), dtype=torch.uint8, device=device)
    names[:, 0:4] = torch.tensor(list(b"SYN."), dtype=torch.uint8, device=device)
    for d in range(9):
        names[:, 4 + d] = ((idx // (10 ** (8 - d))) % 10 + 48).to(torch.uint8)
    out["names"] = names.reshape(-1)
    out["name_off"] = torch.arange(n + 1, device=device, dtype=torch.int64) * W
    return out


def make_core_set(spec, seed=7):
    """Seeded synthetic core set of any size, vectorised: spec = [(length, count)], distinct cores per length, in a
    deterministic (shuffled) order. The reference's patterns.bin is missing from its checkout (.MISSING_LARGE_BLOBS); its
    loader sizes the set for 5-10 M cores (reads.cpp:336, 385). Returns list[str]."""
    rng = np.random.default_rng(seed)
    out = []
    for ln, cnt in spec:
        cnt = int(min(cnt, 4 ** ln))
        have = np.zeros(0, dtype=np.uint64)
        while have.size < cnt:
            need = cnt - have.size
            x = rng.integers(0, 4 ** ln, size=need + need // 4 + 64, dtype=np.uint64)
            have = np.unique(np.concatenate([have, x]))
        have = rng.permutation(have)[:cnt]
        shifts = (2 * (ln - 1 - np.arange(ln))).astype(np.uint64)
        codes = ((have[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
        txt = _BASES[codes]                                   # uint8 [cnt, ln]
        out.extend(txt.view(f"S{ln}").ravel().astype(str).tolist())
    return out

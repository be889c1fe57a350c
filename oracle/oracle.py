"""ctypes front end for the CPU oracle (oracle/liboracle.so) plus the small host-side pieces of the
reference that surround the transform and are needed to compare whole files:

  * output_quality / phred-offset detection  (qualities.cpp:99-104, 177-204)  -> :func:`quality_payload`
  * raw-mode container assembly              (compress.cpp:262-379)           -> :func:`assemble_container`
  * running the real reference CLI           (oracle/_ref/scalce, -T 1)       -> :func:`run_reference_cli`

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg. The product (scalce_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_CLI = os.path.join(HERE, "_ref", "scalce")
MAXBIN = 1 << 30
MAGIC = b"scalce22"


def build(ref: bool | None = None):
    """Compile liboracle.so (always) and oracle/_ref (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isdir("/root/reference")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.orc_submit.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.orc_finish.argtypes = [C.c_void_p]
        L.orc_merge.argtypes = [C.c_void_p]
        L.orc_n_chunks.argtypes = [C.c_void_p]
        L.orc_stream_size.restype = C.c_int64
        L.orc_stream_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_stream_data.restype = C.c_void_p
        L.orc_stream_data.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_unbucketed.argtypes = [C.c_void_p]
        L.orc_n_nodes.restype = C.c_int32
        L.orc_n_nodes.argtypes = [C.c_void_p]
        L.orc_debug.argtypes = [C.c_void_p] * 5
        L.orc_lifetime_count.restype = C.c_uint64
        L.orc_lifetime_count.argtypes = [C.c_void_p, C.c_int32]
        L.orc_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_core_node_id.restype = C.c_int32
        L.orc_core_node_id.argtypes = [C.c_void_p, C.c_int32]
        L.orc_assign.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.orc_parse_fastq.restype = C.c_int64
        L.orc_parse_fastq.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 7
        L.orc_inverse.restype = C.c_int64
        L.orc_inverse.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Sequential (-T 1) restatement of the transform; see scalce_oracle.c."""

    def __init__(self, cores, L1, L2=0, use_names=True, paired=False, use_quals=True, bucket_set_bytes=4 << 30):
        self.cores = [c.encode() if isinstance(c, str) else c for c in cores]
        arr = (C.c_char_p * len(self.cores))(*self.cores)
        self.L1, self.L2, self.paired, self.use_names, self.use_quals = L1, L2, paired, use_names, use_quals
        self.h = lib().orc_create(arr, len(self.cores), L1, L2, int(use_names), int(paired), int(use_quals), bucket_set_bytes)
        self.n = 0

    def submit(self, seq, qual, names, name_off, seq2=None, qual2=None):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = seq.shape[0]
        keep = [seq]
        def prep(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.uint8); keep.append(a); return a
        qual, names, seq2, qual2 = prep(qual), prep(names), prep(seq2), prep(qual2)
        name_off = None if name_off is None else np.ascontiguousarray(name_off, dtype=np.int64)
        lib().orc_submit(self.h, n, _ptr(seq), _ptr(qual), _ptr(names), _ptr(name_off), _ptr(seq2), _ptr(qual2))
        self.n += n

    def assign(self, seq, name_off):
        """Streaming form: bucket id / core / end marker / flush chunk of the next reads, nothing retained (orc_assign).
        Use on a fresh oracle; do not mix with submit()."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = seq.shape[0]
        name_off = None if name_off is None else np.ascontiguousarray(name_off, dtype=np.int64)
        out = [np.empty(n, dtype=np.int32) for _ in range(4)]
        lib().orc_assign(self.h, n, _ptr(seq), _ptr(name_off), *[_ptr(a) for a in out])
        return dict(node_id=out[0], core=out[1], end=out[2], chunk=out[3])

    def inverse(self, stream, seg_core, seg_reads, quals=None, phred=33):
        """decompress.cpp:331-352 for mate 1: (stream 1 without inline headers, bucket table) -> ASCII rows (+ qualities)."""
        stream = np.ascontiguousarray(np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else stream)
        seg_core = np.ascontiguousarray(seg_core, dtype=np.int32); seg_reads = np.ascontiguousarray(seg_reads, dtype=np.int64)
        n = int(seg_reads.sum())
        q = None if quals is None else np.ascontiguousarray(np.frombuffer(quals, dtype=np.uint8) if isinstance(quals, (bytes, bytearray)) else quals)
        seq = np.empty((max(n, 1), self.L1), dtype=np.uint8)
        qo = None if q is None else np.empty((max(n, 1), self.L1), dtype=np.uint8)
        got = lib().orc_inverse(self.h, _ptr(stream), _ptr(seg_core), _ptr(seg_reads), len(seg_core), _ptr(q), phred, _ptr(seq), _ptr(qo))
        assert got == n
        return seq[:n], (None if qo is None else qo[:n])

    def finish(self):
        lib().orc_finish(self.h)
        lib().orc_merge(self.h)

    @property
    def n_chunks(self):
        return lib().orc_n_chunks(self.h)

    def stream(self, k, chunk=-1) -> bytes:
        """Stream k (0 names,1 reads,2 quals,3 meta,4 reads2,5 quals2) of a flush chunk, or merged (-1)."""
        n = lib().orc_stream_size(self.h, chunk, k)
        if n == 0:
            return b""
        return C.string_at(lib().orc_stream_data(self.h, chunk, k), n)

    def debug(self):
        out = [np.empty(self.n, dtype=np.int32) for _ in range(4)]
        lib().orc_debug(self.h, *[_ptr(a) for a in out])
        return dict(node_id=out[0], core=out[1], end=out[2], chunk=out[3])

    def lifetime_count(self, core):
        return lib().orc_lifetime_count(self.h, core)

    def candidates(self, text: np.ndarray, cap=64):
        cc = np.empty(cap, dtype=np.int32); cp = np.empty(cap, dtype=np.int32); lv = np.zeros(1, dtype=np.int32)
        text = np.ascontiguousarray(text, dtype=np.uint8)
        n = lib().orc_candidates(self.h, _ptr(text), text.size, cap, _ptr(cc), _ptr(cp), _ptr(lv))
        return int(lv[0]), cc[:min(n, cap)].copy(), cp[:min(n, cap)].copy(), n

    def core_node_id(self, core):
        return lib().orc_core_node_id(self.h, core)

    @property
    def unbucketed(self):
        return lib().orc_unbucketed(self.h)

    @property
    def n_nodes(self):
        return lib().orc_n_nodes(self.h)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# host-side neighbours of the transform, restated for whole-file comparison
# ------------------------------------------------------------------------------------------------
def detect_phred_offset(qual: np.ndarray, sample=100000) -> int:
    """qualities.cpp:99-104: 33 if any quality char in [33,64) appears in the first `sample` records."""
    q = qual[:sample]
    return 33 if ((q >= 33) & (q < 64)).any() else 64


def quality_payload(qual: np.ndarray, seq: np.ndarray, offset: int) -> np.ndarray:
    """output_quality, qualities.cpp:177-204 at lossy percentage 0: q-offset, 0 where the base is 'N'."""
    out = (qual.astype(np.int16) - offset).astype(np.uint8)
    out[seq == ord("N")] = 0
    return out


def assemble_container(meta: bytes, names: bytes, reads: bytes, quals: bytes, cores, L, phred_offset, *,
                       use_names=True, library=b"", paired=False, mate=0, no_ac=1, reads2=b"", quals2=b"", L2=0):
    """Raw-mode (-c no -A) .scalcen/.scalcer/.scalceq bytes for one mate from merged streams
    (combine_and_compress_with_split, compress.cpp:262-379). Returns (n_bytes, r_bytes, q_bytes)."""
    nlen = 3 + 2 * int(paired)
    rsz = 8 + 8 * nlen
    sz_meta = 2 if L > 255 else 1
    fn = bytearray(MAGIC + bytes([1 if use_names else 0]))
    if not use_names:
        fn += struct.pack("<q", 0) + library
    fr = bytearray(MAGIC + struct.pack("<i", no_ac) + struct.pack("<i", L2 if mate else L))
    fq = bytearray(MAGIC + struct.pack("<q", phred_offset))
    pn = pr = pq = 0
    src_r, src_q = (reads2, quals2) if mate else (reads, quals)
    for o in range(0, len(meta), rsz):
        _id, core = struct.unpack_from("<ii", meta, o)
        lens = struct.unpack_from("<%dq" % nlen, meta, o + 8)
        lN, lR, lQ = lens[0], lens[1], lens[2]
        if mate:
            lR, lQ = lens[3], lens[4]
        else:
            clen = 0 if core == MAXBIN - 1 else len(cores[core])
            size = lR // ((L - clen + 3) // 4 + sz_meta)
            fr += struct.pack("<iq", core, size)
        fr += src_r[pr:pr + lR]; pr += lR
        fq += src_q[pq:pq + lQ]; pq += lQ
        if use_names:
            fn += names[pn:pn + lN]; pn += lN
    return bytes(fn), bytes(fr), bytes(fq)


def parse_fastq(text: bytes, L, phred=33, stats=None):
    """The reference's host front end restated (compress.cpp:614-671, names.cpp:48-62, qualities.cpp:177-204): FASTQ text ->
    (seq [n, L], qual payload [n, L], names, name_off). stats = dict(freq3, freq4, prev) carried across calls (None: off)."""
    t = np.frombuffer(text, dtype=np.uint8)
    cap = t.size // (2 * L + 4) + 2
    seq = np.empty((cap, L), dtype=np.uint8); qual = np.empty((cap, L), dtype=np.uint8)
    names = np.empty(t.size, dtype=np.uint8); off = np.zeros(cap + 1, dtype=np.int64)
    f3 = f4 = pv = None
    if stats is not None:
        f3, f4, pv = stats["freq3"], stats["freq4"], stats["prev"]
    n = lib().orc_parse_fastq(_ptr(t), t.size, L, phred, _ptr(seq), _ptr(qual), _ptr(names), _ptr(off), _ptr(f3), _ptr(f4), _ptr(pv))
    if n < 0:
        raise ValueError("malformed FASTQ")
    return seq[:n], qual[:n], names[:off[n]], off[:n + 1]


def new_quality_stats():
    return dict(freq3=np.zeros(80 * 80, dtype=np.uint64), freq4=np.zeros(80 * 80 * 80, dtype=np.uint64), prev=np.array([500, 500], dtype=np.uint32))


def segments_from_meta(meta: bytes, cores, L, paired=False):
    """(core index, reads) of every bucket record, as combine_and_compress_with_split derives them (compress.cpp:364-379)."""
    nlen = 3 + 2 * int(paired)
    rsz = 8 + 8 * nlen
    sz_meta = 2 if L > 255 else 1
    sc, sr = [], []
    for o in range(0, len(meta), rsz):
        _id, core = struct.unpack_from("<ii", meta, o)
        lR = struct.unpack_from("<q", meta, o + 16)[0]
        clen = 0 if core == MAXBIN - 1 else len(cores[core])
        sc.append(core); sr.append(lR // ((L - clen + 3) // 4 + sz_meta))
    return np.array(sc, dtype=np.int32), np.array(sr, dtype=np.int64)


def split_reads_container(body: bytes, cores, L):
    """.scalcer after its 16 header bytes -> (stream 1 without the inline records, seg_core, seg_reads): the walk the
    decompressor does over the bucket headers (decompress.cpp:262-272)."""
    sz_meta = 2 if L > 255 else 1
    pos, parts, sc, sr = 0, [], [], []
    while pos < len(body):
        core, cnt = struct.unpack_from("<iq", body, pos)
        pos += 12
        clen = 0 if core == MAXBIN - 1 else len(cores[core])
        nb = cnt * ((L - clen + 3) // 4 + sz_meta)
        parts.append(body[pos:pos + nb]); pos += nb
        sc.append(core); sr.append(cnt)
    return b"".join(parts), np.array(sc, dtype=np.int32), np.array(sr, dtype=np.int64)


def run_reference_cli(fastq1, out_prefix, cores_txt=None, *, paired=False, bucket="4G", raw=True, no_names=None,
                      tmpdir=None, extra=()):
    """Run the unmodified reference at -T 1 (the only deterministic mode, SURVEY.md preamble 3)."""
    if not os.path.exists(REF_CLI):
        raise FileNotFoundError(REF_CLI)
    cmd = [REF_CLI, fastq1, "-T", "1", "-o", out_prefix, "-B", bucket]
    if cores_txt:
        cmd += ["-P", cores_txt]
    if raw:
        cmd += ["-c", "no", "-A"]
    if paired:
        cmd += ["-r"]
    if no_names is not None:
        cmd += ["-n", no_names]
    if tmpdir:
        cmd += ["-t", tmpdir]
    cmd += list(extra)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=os.path.dirname(out_prefix) or ".")
    if r.returncode != 0:
        raise RuntimeError("reference CLI failed: %s\n%s" % (" ".join(cmd), r.stderr.decode(errors="replace")))
    return r


def run_reference_decompress(scalcen_path, out_prefix, cores_txt=None, paired=False):
    cmd = [REF_CLI, scalcen_path, "-d", "-T", "1", "-o", out_prefix]
    if cores_txt:
        cmd += ["-P", cores_txt]
    if paired:
        cmd += ["-r"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=os.path.dirname(out_prefix) or ".")
    if r.returncode != 0:
        raise RuntimeError("reference decompress failed: %s\n%s" % (" ".join(cmd), r.stderr.decode(errors="replace")))
    return r

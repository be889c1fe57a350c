"""ctypes binding of the C ABI (include/scalce_b200.h) and a small host-side wrapper.

The product path is libscalce_b200.so (hand-written sm_100a CUDA). There is no CPU fallback: if
the library cannot be loaded or no B200-class device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCB_LIB") or os.path.join(HERE, "libscalce_b200.so")   # SCB_LIB: an alternative build (tuning experiments)

N_STREAMS = 6
S_NAMES, S_READS, S_QUALS, S_META, S_READS2, S_QUALS2 = range(6)
ROOT_ID = (1 << 30) - 1

EXPORTS = [
    "scb_abi_version", "scb_last_error", "scb_table_dryrun", "scb_create", "scb_create_from_file", "scb_table_info", "scb_core",
    "scb_submit", "scb_flush", "scb_flush_closed", "scb_copy_stream", "scb_copy_debug", "scb_unbucketed", "scb_lifetime_count",
    "scb_kernel_launches", "scb_stage_ms", "scb_resolve_rounds", "scb_resolve_engine", "scb_device_bytes", "scb_assemble_reads", "scb_copy_assembled", "scb_inverse_reads", "scb_submit_fastq", "scb_quality_stats", "scb_reset_counts", "scb_destroy",
    "scb_set_stream", "scb_shard_info", "scb_shard_scan", "scb_shard_sizes", "scb_shard_resolve_local", "scb_shard_resolve_round",
    "scb_shard_finalize", "scb_shard_bucket_hist", "scb_shard_pack", "scb_shard_import", "scb_shard_finish", "scb_shard_last_ms",
    "scb_shard_partition", "scb_shard_recv_reserve", "scb_shard_send", "scb_shard_send_wait", "scb_shard_finish_sort", 
    "scb_ipc_export", "scb_ipc_open", "scb_ipc_close", "scb_shard_flush", "scb_shard_flush_stats", "scb_shard_n_local",
    "scb_shard_chunk_layout", "scb_shard_partition_chunks", "scb_shard_split_mode", "scb_shard_chunk_owners", "scb_shard_flush_wall",
]


class ScbConfig(C.Structure):
    _fields_ = [("read_length", C.c_int32 * 2), ("use_names", C.c_int32), ("paired", C.c_int32), ("use_quals", C.c_int32),
                ("device", C.c_int32), ("bucket_set_bytes", C.c_uint64), ("emit_merged", C.c_int32), ("reserved", C.c_int32)]


class ScbBatch(C.Structure):
    _fields_ = [("n", C.c_int64), ("seq1", C.c_void_p), ("qual1", C.c_void_p), ("names", C.c_void_p), ("name_off", C.c_void_p),
                ("seq2", C.c_void_p), ("qual2", C.c_void_p), ("location", C.c_int32), ("reserved", C.c_int32)]


class ScbResult(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_chunks", C.c_int32), ("n_buckets_nonempty", C.c_int32),
                ("data", C.c_void_p * N_STREAMS), ("chunk_off", C.POINTER(C.c_int64) * N_STREAMS),
                ("merged", C.c_void_p * N_STREAMS), ("merged_size", C.c_int64 * N_STREAMS),
                ("bucket_id", C.c_void_p), ("core_idx", C.c_void_p), ("end", C.c_void_p), ("chunk", C.c_void_p),
                ("perm", C.c_void_p), ("device_ms", C.c_float)]


class ScbShardXfer(C.Structure):
    _fields_ = [("n", C.c_int64), ("name_bytes", C.c_int64), ("aux", C.c_void_p), ("packed", C.c_void_p), ("qual1", C.c_void_p),
                ("names", C.c_void_p), ("seq2", C.c_void_p), ("qual2", C.c_void_p), ("cnt_reads", C.POINTER(C.c_int64)),
                ("cnt_name_bytes", C.POINTER(C.c_int64)), ("packed_row_bytes", C.c_int32), ("reserved", C.c_int32)]


class ScbShardPeer(C.Structure):
    _fields_ = [("aux", C.c_void_p), ("packed", C.c_void_p), ("qual1", C.c_void_p), ("names", C.c_void_p), ("seq2", C.c_void_p),
                ("qual2", C.c_void_p), ("row_off", C.c_int64), ("name_off", C.c_int64)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class ScbComm(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_ranks", C.c_int32), ("same_process", C.c_int32), ("reserved", C.c_int32), ("ctx", C.c_void_p),
                ("allgather", ALLGATHER_FN), ("barrier", BARRIER_FN)]


_lib = None


def load_library(path: str | None = None):
    """Loads the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SCB_LIB_PATH") or LIB_PATH     # SCB_LIB_PATH: a differently built library for A/B runs of compile-time choices
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `python -m scalce_b200.build` (nvcc, sm_100a). "
                           "scalce_b200 has no CPU fallback.")
    L = C.CDLL(p)
    L.scb_abi_version.restype = C.c_int
    L.scb_last_error.restype = C.c_char_p
    L.scb_create.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.POINTER(ScbConfig), C.POINTER(C.c_void_p)]
    L.scb_create_from_file.argtypes = [C.c_char_p, C.POINTER(ScbConfig), C.POINTER(C.c_void_p)]
    L.scb_table_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
    L.scb_table_dryrun.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
    L.scb_core.restype = C.c_char_p
    L.scb_core.argtypes = [C.c_void_p, C.c_int32]
    L.scb_submit.argtypes = [C.c_void_p, C.POINTER(ScbBatch)]
    L.scb_flush.argtypes = [C.c_void_p, C.POINTER(ScbResult)]
    L.scb_flush_closed.argtypes = [C.c_void_p, C.POINTER(ScbResult)]
    L.scb_copy_stream.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
    L.scb_copy_debug.argtypes = [C.c_void_p] * 6
    L.scb_unbucketed.restype = C.c_int64
    L.scb_unbucketed.argtypes = [C.c_void_p]
    L.scb_lifetime_count.restype = C.c_int64
    L.scb_lifetime_count.argtypes = [C.c_void_p, C.c_int32]
    L.scb_kernel_launches.restype = C.c_int64
    L.scb_kernel_launches.argtypes = [C.c_void_p]
    L.scb_stage_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int32]
    L.scb_resolve_rounds.argtypes = [C.c_void_p]
    L.scb_resolve_engine.argtypes = [C.c_void_p]
    L.scb_submit_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    L.scb_quality_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.scb_assemble_reads.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.scb_copy_assembled.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.scb_inverse_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    L.scb_device_bytes.restype = C.c_int64
    L.scb_device_bytes.argtypes = [C.c_void_p]
    L.scb_reset_counts.argtypes = [C.c_void_p]
    L.scb_destroy.argtypes = [C.c_void_p]
    L.scb_set_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.scb_shard_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.scb_shard_scan.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.scb_shard_sizes.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    L.scb_shard_resolve_local.argtypes = [C.c_void_p, C.c_void_p]
    L.scb_shard_resolve_round.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    L.scb_shard_finalize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.scb_shard_bucket_hist.argtypes = [C.c_void_p, C.c_void_p]
    L.scb_shard_pack.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(ScbShardXfer)]
    L.scb_shard_import.argtypes = [C.c_void_p, C.POINTER(ScbShardXfer), C.c_int32]
    L.scb_shard_finish.argtypes = [C.c_void_p, C.POINTER(ScbResult)]
    L.scb_shard_partition.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.POINTER(ScbShardXfer)]
    L.scb_shard_recv_reserve.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    L.scb_shard_send.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ScbShardPeer), C.c_int32, C.c_int32]
    L.scb_shard_send_wait.argtypes = [C.c_void_p]
    L.scb_shard_finish_sort.argtypes = [C.c_void_p]
    L.scb_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.scb_ipc_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.scb_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    L.scb_shard_flush.argtypes = [C.c_void_p, C.POINTER(ScbComm), C.POINTER(ScbResult)]
    L.scb_shard_n_local.restype = C.c_int64
    L.scb_shard_n_local.argtypes = [C.c_void_p]
    L.scb_shard_flush_stats.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]
    L.scb_shard_last_ms.restype = C.c_float
    L.scb_shard_last_ms.argtypes = [C.c_void_p]
    L.scb_shard_chunk_layout.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.scb_shard_partition_chunks.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(ScbShardXfer)]
    L.scb_shard_split_mode.argtypes = [C.c_void_p]
    L.scb_shard_flush_wall.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int32]
    L.scb_shard_chunk_owners.argtypes = [C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    if path is None:
        _lib = L
    return L


class ScbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"scalce_b200 error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise ScbError(rc, load_library().scb_last_error().decode(errors="replace"))


def table_dryrun(cores):
    """Host-only compile of a core set (no device needed): shape of the automaton and per-core node ids."""
    enc = [c.encode() if isinstance(c, str) else c for c in cores]
    arr = (C.c_char_p * len(enc))(*enc)
    ns, nb, rp = C.c_int32(), C.c_int32(), C.c_int32()
    ids = np.empty(len(enc), dtype=np.int32)
    _check(load_library().scb_table_dryrun(arr, len(enc), C.byref(ns), C.byref(nb), C.byref(rp), ids.ctypes.data_as(C.c_void_p)))
    return dict(n_states=ns.value, n_buckets=nb.value, root_order_pos=rp.value, core_node_id=ids)


class FlushResult:
    """Host-side view of one scb_flush."""

    def __init__(self, tr, res: ScbResult):
        self._tr = tr
        self.n_reads = res.n_reads
        self.n_chunks = res.n_chunks
        self.n_buckets_nonempty = res.n_buckets_nonempty
        self.device_ms = res.device_ms
        self.chunk_off = [[res.chunk_off[k][c] for c in range(res.n_chunks + 1)] for k in range(N_STREAMS)]
        self.merged_size = [res.merged_size[k] for k in range(N_STREAMS)]
        self.data_ptr = [res.data[k] for k in range(N_STREAMS)]
        self.merged_ptr = [res.merged[k] for k in range(N_STREAMS)]
        # device pointers of the per-read arrays (int32, input order) and of perm (uint32, output order)
        self.array_ptr = dict(node_id=res.bucket_id, core=res.core_idx, end=res.end, chunk=res.chunk, perm=res.perm)

    def stream(self, k, chunk=-1) -> bytes:
        """Bytes of stream k for a flush chunk (the t_%03d_k.tmp contents) or merged (chunk=-1)."""
        if chunk < 0:
            n = self.merged_size[k]
        else:
            n = self.chunk_off[k][chunk + 1] - self.chunk_off[k][chunk]
        buf = np.empty(max(n, 1), dtype=np.uint8)
        _check(load_library().scb_copy_stream(self._tr._h, k, chunk, buf.ctypes.data_as(C.c_void_p), n))
        return buf[:n].tobytes()

    def torch_view(self, what, n=None):
        """Zero-copy torch view of a device-resident result: 'perm' / 'node_id' / 'core' / 'end' / 'chunk' (int32; perm's
        uint32 bit patterns) or a stream number (uint8, all chunks). Valid until the next submit / flush."""
        import torch

        class _P:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        dev = torch.device("cuda", self._tr.cfg.device)
        if isinstance(what, int):
            nb = self.chunk_off[what][-1]
            return torch.as_tensor(_P(self.data_ptr[what], nb), device=dev) if nb else torch.empty(0, dtype=torch.uint8, device=dev)
        cnt = self.n_reads if n is None else n
        if cnt == 0:
            return torch.empty(0, dtype=torch.int32, device=dev)
        return torch.as_tensor(_P(self.array_ptr[what], cnt * 4), device=dev).view(torch.int32)

    def debug(self, n_local=None):
        """Per-read arrays in input order (of the rank's own shard after a sharded run: pass n_local) and perm."""
        n = self.n_reads if n_local is None else n_local
        arrs = [np.empty(n, dtype=np.int32) for _ in range(4)] + [np.empty(self.n_reads, dtype=np.uint32)]
        _check(load_library().scb_copy_debug(self._tr._h, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
        return dict(node_id=arrs[0], core=arrs[1], end=arrs[2], chunk=arrs[3], perm=arrs[4])


class BoostTransform:
    """Host-side mirror of the reference's seam for this path (reads.h:87-94 in batch form).

    create  <- read_patterns[_from_file] + prepare_aho_automata
    submit  <- per-read aho_search / output_read / aho_trie_bucket + size accounting
    flush   <- dump_trie / aho_output (+ merge() when emit_merged)
    """

    def __init__(self, cores=None, L1=0, L2=0, *, core_file=None, use_names=True, paired=False, use_quals=True,
                 bucket_set_bytes=4 << 30, device=0, emit_merged=True):
        L = load_library()
        cfg = ScbConfig()
        cfg.read_length[0] = L1
        cfg.read_length[1] = L2 if paired else 0
        cfg.use_names, cfg.paired, cfg.use_quals = int(use_names), int(paired), int(use_quals)
        cfg.device, cfg.bucket_set_bytes, cfg.emit_merged = device, bucket_set_bytes, int(emit_merged)
        self.cfg = cfg
        h = C.c_void_p()
        if core_file is not None:
            _check(L.scb_create_from_file(os.fsencode(core_file), C.byref(cfg), C.byref(h)))
        else:
            enc = [c.encode() if isinstance(c, str) else c for c in cores]
            arr = (C.c_char_p * len(enc))(*enc)
            _check(L.scb_create(arr, len(enc), C.byref(cfg), C.byref(h)))
        self._h = h
        self._keep = []

    # -- info ---------------------------------------------------------------------------------
    def table_info(self):
        v = [C.c_int32() for _ in range(4)]
        _check(load_library().scb_table_info(self._h, *[C.byref(x) for x in v]))
        return dict(n_cores=v[0].value, n_states=v[1].value, n_buckets=v[2].value, smem_resident=bool(v[3].value))

    def core(self, idx):
        s = load_library().scb_core(self._h, idx)
        return None if s is None else s.decode()

    # -- data ---------------------------------------------------------------------------------
    def submit(self, seq, qual=None, names=None, name_off=None, seq2=None, qual2=None):
        """Host numpy arrays (uint8 [n,L]; names uint8, name_off int64 [n+1])."""
        def prep(a, dt=np.uint8):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=dt)
            return a, a.ctypes.data_as(C.c_void_p)
        keep = []
        b = ScbBatch()
        seq, b.seq1 = prep(seq); b.n = seq.shape[0]
        qual, b.qual1 = prep(qual); names, b.names = prep(names); name_off, b.name_off = prep(name_off, np.int64)
        seq2, b.seq2 = prep(seq2); qual2, b.qual2 = prep(qual2)
        keep += [seq, qual, names, name_off, seq2, qual2]
        b.location = 0
        _check(load_library().scb_submit(self._h, C.byref(b)))

    def submit_device(self, n, seq, qual=None, names=None, name_off=None, seq2=None, qual2=None):
        """Raw device pointers (ints), e.g. torch tensors' data_ptr(); buffers must outlive flush()."""
        b = ScbBatch()
        b.n = n
        b.seq1, b.qual1, b.names, b.name_off, b.seq2, b.qual2 = seq, qual, names, name_off, seq2, qual2
        b.location = 1
        _check(load_library().scb_submit(self._h, C.byref(b)))

    def flush(self) -> FlushResult:
        res = ScbResult()
        _check(load_library().scb_flush(self._h, C.byref(res)))
        return FlushResult(self, res)

    def submit_fastq(self, text1, text2=None, phred_offset=(33, 33)):
        """(f2) FASTQ text (bytes / uint8 arrays on the host) parsed on the device into one pending batch. Returns the record count."""
        L = load_library()
        a1 = np.frombuffer(text1, dtype=np.uint8) if isinstance(text1, (bytes, bytearray)) else np.ascontiguousarray(text1, dtype=np.uint8)
        a2 = None if text2 is None else (np.frombuffer(text2, dtype=np.uint8) if isinstance(text2, (bytes, bytearray)) else np.ascontiguousarray(text2, dtype=np.uint8))
        ph = (C.c_int32 * 2)(*phred_offset)
        n = C.c_int64()
        _check(L.scb_submit_fastq(self._h, a1.ctypes.data_as(C.c_void_p), a1.size, None if a2 is None else a2.ctypes.data_as(C.c_void_p),
                                  0 if a2 is None else a2.size, 0, ph, C.byref(n)))
        return n.value

    def quality_stats(self, mate=0):
        """ac_freq3 [80, 80] and ac_freq4 [80, 80, 80] (uint64) as output_quality accumulates them in input order."""
        f3 = np.zeros(80 * 80, dtype=np.uint64); f4 = np.zeros(80 * 80 * 80, dtype=np.uint64)
        _check(load_library().scb_quality_stats(self._h, mate, f3.ctypes.data_as(C.c_void_p), f4.ctypes.data_as(C.c_void_p)))
        return f3.reshape(80, 80), f4.reshape(80, 80, 80)

    def assemble_reads(self, chunk=-1):
        """(f3) .scalcer body of mate 1 (bucket records + packed reads, compress.cpp:345-384) from the last flush, assembled on
        the device. Returns (body bytes, seg_core int32 [n_seg], seg_reads int64 [n_seg])."""
        L = load_library()
        nb, ns = C.c_int64(), C.c_int64()
        _check(L.scb_assemble_reads(self._h, chunk, None, C.byref(nb), C.byref(ns)))
        body = np.empty(max(nb.value, 1), dtype=np.uint8)
        sc = np.empty(max(ns.value, 1), dtype=np.int32); sr = np.empty(max(ns.value, 1), dtype=np.int64)
        _check(L.scb_copy_assembled(self._h, body.ctypes.data_as(C.c_void_p), nb.value, sc.ctypes.data_as(C.c_void_p), sr.ctypes.data_as(C.c_void_p)))
        return body[:nb.value].tobytes(), sc[:ns.value], sr[:ns.value]

    def inverse_reads(self, stream, seg_core, seg_reads, quals=None, mate=0, phred_offset=33):
        """(f4) decompress-side inverse on the device (decompress.cpp:331-352), host buffers in and out.
        Returns (seq uint8 [n, L], qual uint8 [n, L] or None)."""
        L = load_library()
        Lr = self.cfg.read_length[1 if mate else 0]
        stream = np.ascontiguousarray(np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else stream, dtype=np.uint8)
        seg_core = np.ascontiguousarray(seg_core, dtype=np.int32); seg_reads = np.ascontiguousarray(seg_reads, dtype=np.int64)
        n = int(seg_reads.sum())
        seq = np.empty((max(n, 1), Lr), dtype=np.uint8)
        q_in = None if quals is None else np.ascontiguousarray(np.frombuffer(quals, dtype=np.uint8) if isinstance(quals, (bytes, bytearray)) else quals, dtype=np.uint8)
        q_out = None if quals is None else np.empty((max(n, 1), Lr), dtype=np.uint8)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        nn = C.c_int64()
        _check(L.scb_inverse_reads(self._h, p(stream), p(seg_core), p(seg_reads), len(seg_core), p(q_in), mate, phred_offset, 0, p(seq), p(q_out), C.byref(nn)))
        assert nn.value == n
        return seq[:n], (None if q_out is None else q_out[:n])

    def flush_closed(self) -> FlushResult:
        """Streaming flush: only the complete flush chunks; the open chunk's reads stay pending (scb_flush_closed)."""
        res = ScbResult()
        _check(load_library().scb_flush_closed(self._h, C.byref(res)))
        return FlushResult(self, res)

    @property
    def unbucketed(self):
        return load_library().scb_unbucketed(self._h)

    def lifetime_count(self, core_idx):
        return load_library().scb_lifetime_count(self._h, core_idx)

    STAGES = ("scan", "resolve", "chunks", "sort", "ties", "emit", "merged", "arrays")

    def stage_ms(self):
        buf = (C.c_float * 8)()
        load_library().scb_stage_ms(self._h, buf, 8)
        return dict(zip(self.STAGES, [float(x) for x in buf]))

    def reset_counts(self):
        _check(load_library().scb_reset_counts(self._h))

    @property
    def resolve_rounds(self):
        return load_library().scb_resolve_rounds(self._h)

    def n_local_last(self):
        """Reads of this rank's own input shard in the last sharded flush (the per-read arrays refer to them)."""
        return int(load_library().scb_shard_n_local(self._h))

    @property
    def device_bytes(self):
        """High-water mark of the flush workspace in bytes."""
        return load_library().scb_device_bytes(self._h)

    @property
    def resolve_engine(self):
        """0 dense (shared-memory population rows), 1 sparse (bucket-major candidate lists), 2 sequential."""
        return load_library().scb_resolve_engine(self._h)

    @property
    def kernel_launches(self):
        return load_library().scb_kernel_launches(self._h)

    def close(self):
        if getattr(self, "_h", None):
            load_library().scb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

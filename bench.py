#!/usr/bin/env python3
"""bench.py - throughput of the boosting transform (core scan + bucket + reorder) on B200.

A step = one pass of the hot path (scb_submit + scb_flush through the C ABI) over one batch of
synthetic fixed-length reads. Workload at 1 GPU = BASELINE.json configs[1]: 50M x 150 bp
single-end; under torchrun every rank holds the same per-GPU workload (weak scaling), the global
input is the ranks' batches concatenated in rank order, and the sharded run (joint
exact tie-break, bucket-range exchange over NVLink peer memory) produces the single-GPU order of it.

  value  reads/s with inputs resident in HBM when the timed region starts (CUDA events on the
         stream the library runs on, max over ranks)
  e2e    reads/s through the same C ABI with HOST (pinned) buffers: H2D of all inputs and D2H
         of every output stream inside the timed region
  --impl reference : the reference's own CPU transform (oracle/_ref, unmodified objects)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEADLINE_SPEC = [(8, 1024), (9, 512), (10, 256), (11, 128), (12, 128)]  # 2048 synthetic cores, seed 7
BIG_SPEC_FRAC = [(10, 0.1), (11, 0.2), (12, 0.4), (13, 0.2), (14, 0.1)]  # --cores N >= 100000: N cores of 10-14 bases (the reference sizes patterns[] for 5-10 M, reads.cpp:336, 385)
CORE_SEED = 7
NAME_BYTES = 13  # "SYN.%09d"

# BASELINE.json configs[1..4]. reads = reads (pairs) per GPU; total = reads of the whole job, split over the GPUs.
CONFIGS = {
    "c2": dict(reads=50_000_000, L=150, paired=False, L2=0, high_entropy=False, scaling="weak",
               desc="synthetic {n}M x 150bp single-end FASTQ per GPU (BASELINE configs[1])"),
    "c3": dict(reads=25_000_000, L=150, paired=True, L2=150, high_entropy=False, scaling="weak",
               desc="synthetic {n}M x 150bp paired-end FASTQ per GPU, mates kept in sync through the reorder (BASELINE configs[2]: 200M pairs over 8 GPUs)"),
    "c4": dict(total=500_000_000, L=36, paired=False, L2=0, high_entropy=False, scaling="strong",
               desc="synthetic 500M x 36bp single-end FASTQ in total = {n}M per GPU (BASELINE configs[3])"),
    "c5": dict(reads=12_500_000, L=250, paired=False, L2=0, high_entropy=True, scaling="weak",
               desc="synthetic {n}M x 250bp single-end FASTQ per GPU, high-entropy qualities (BASELINE configs[4]: 100M reads over 8 GPUs)"),
}


# DRAM bytes per read of each stage's kernels at the headline shape: dram__bytes_read.sum + dram__bytes_write.sum, summed over
# the stage's launches of one ncu pass over a headline step (profiles/r02_final_launches_summary.txt)
NCU_TRAFFIC_PER_READ = {"emit": 736, "resolve": 108, "scan": 227, "sort": 226, "ties": 15, "chunks": 28, "arrays": 16}
NCU_TRAFFIC_SOURCE = "ncu launch list of the headline step with this round's kernels (profiles/r02_final_launches_summary.txt): dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's launches"
STAGE_KERNELS = {"emit": "gather_rows16_k + emit_reads_fast_k + emit_names_st_k + emit_off_reduce_k / emit_off_apply_k (metadata gather + offset scans)",
                 "resolve": "resolve_dense_k + resolve_finalize_k (dense) or sp_flags_k / sp_tilescan_k / sp_counts_k / sp_decide_k per round (sparse)",
                 "scan": "scan_smem2_k (table in shared memory) or scan_big_k (table in global memory / L2)",
                 "sort": "build_keys_pk_k + n x (sort_hist_k, sort_scatter_k)", "exchange_rows": "gather_rows16_to_k (peer stores)"}


def algorithmic_bytes_per_read(L, name_len, mean_core, paired=False, L2=0):
    """read seq L + qual L + name; write name + packed + end marker + qual L."""
    b = 3 * L + 2 * (name_len + 1) + (L - mean_core + 3) // 4 + (2 if L > 255 else 1)
    if paired:
        b += 3 * L2 + (L2 + 3) // 4
    return b


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a host thread every ~10 ms
    (a 50M-read step is ~35 ms), `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines = []
        self.proc = None
        self.nvml = None
        self.samples = []      # (sm_mhz, reasons bitmask)
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML orders devices by PCI bus id; honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")) and index < len(vis.split(",")):
                phys = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # both calls of a sample must work here, otherwise nvidia-smi does the sampling
            self._reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            try:
                self._reasons(self.h)
            except Exception:
                self._reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                self._reasons(self.h)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
        except Exception:
            self.proc = None

    def start(self):
        """Sampling starts here; construction (NVML initialisation: ~0.1 s) belongs before the barrier that opens the timed region -
        at N > 1 a rank 0 that enters its first timed step late makes every other rank wait inside that step."""
        if self.nvml is not None or self.proc is not None:
            self.samples.clear(); self.lines.clear()
            self.t.start()
        return self

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.samples.append((mhz, int(self._reasons(self.h))))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
            nv = self.nvml
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = [x[0] for x in self.samples]
            reasons = sorted(nm for nm, b in bits.items() if any(x[1] & b for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm), "source": "nvml, 10 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}


def bench_cores(n_cores):
    """The core set of a bench line and its description: 2048 (default) = the shared-memory resident headline set; >= 100000 =
    N cores of 10-14 bases (automaton in global memory / L2, sparse tie-break engine)."""
    if n_cores in (0, 2048):
        c = headline_cores()
        return c, f"{len(c)} synthetic cores 8-12bp seed {CORE_SEED} (automaton resident in shared memory)", 8
    from scalce_b200 import synth
    spec = [(ln, max(1, int(round(n_cores * fr)))) for ln, fr in BIG_SPEC_FRAC]
    c = synth.make_core_set(spec, seed=CORE_SEED)
    return c, f"{len(c)} synthetic cores 10-14bp seed {CORE_SEED} (automaton in global memory, served by L2)", 12


def headline_cores():
    # same generator as oracle/gen_cores.py, duplicated here because the product bench must not
    # import oracle/ outside the cpu_baseline / reference legs
    rng = np.random.default_rng(CORE_SEED)
    out = []
    for ln, cnt in HEADLINE_SPEC:
        seen = set()
        while len(seen) < cnt:
            codes = rng.integers(0, 4, size=((cnt - len(seen)) * 2 + 8, ln), dtype=np.uint8)
            for row in codes:
                s = "".join("ACGT"[c] for c in row)
                if s not in seen:
                    seen.add(s); out.append(s)
                    if len(seen) == cnt:
                        break
    return out


_REF = {}   # the harness library and the core set it was initialised with (once per process: the reference keeps its pattern
            # counter in a file-static and has no way to unload a core set, reads.cpp:52)


def run_reference_harness(cores, seq, qual, names, name_off, L, threads=1, seq2=None, qual2=None, L2=0):
    """Times the unmodified reference objects (oracle/_ref/libref_harness.so) or, if that was not
    built, the oracle port, on host arrays. threads > 1 runs the harness' copy of the reference's -T loop
    (same locks as compress.cpp thread(); output not deterministic, never compared). The port is single-threaded.
    Bucket populations carry over from one call to the next (the reference never resets bin_size, reads.h:82).
    Returns (reads_per_s, kind, seconds, threads_used)."""
    n = seq.shape[0]
    paired = seq2 is not None
    harness = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    d = tempfile.mkdtemp(prefix="scb_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        if os.path.exists(harness):
            key = (L, L2, paired, len(cores), hash(tuple(cores[:64])))
            if "H" not in _REF:
                H = C.CDLL(harness)
                H.refh_init.restype = C.c_double
                H.refh_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
                H.refh_run.restype = C.c_double
                H.refh_run.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
                with open(os.path.join(d, "cores.txt"), "w") as f:
                    f.write("\n".join(cores) + "\n")
                # the harness prints the reference's own LOG lines to stderr
                H.refh_init(os.path.join(d, "cores.txt").encode(), L, L2, 1 if paired else 0, 1, 4 << 30)
                _REF["H"], _REF["key"] = H, key
            if _REF["key"] != key:
                raise RuntimeError("the reference harness serves one core set / read shape per process")
            H = _REF["H"]
            p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
            nch = C.c_int()
            if threads > 1 and hasattr(H, "refh_run_mt"):
                H.refh_run_mt.restype = C.c_double
                H.refh_run_mt.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_void_p, C.c_int]
                secs = H.refh_run_mt(n, p(seq), p(qual), p(names), p(name_off), p(seq2), p(qual2), 33, d.encode(), C.byref(nch), threads)
                return n / secs, "reference", secs, threads
            secs = H.refh_run(n, p(seq), p(qual), p(names), p(name_off), p(seq2), p(qual2), 33, d.encode(), C.byref(nch), None, None)
            return n / secs, "reference", secs, 1
        from oracle import oracle as orc
        q1 = orc.quality_payload(qual, seq, 33)
        q2 = orc.quality_payload(qual2, seq2, 33) if paired else None
        t0 = time.perf_counter()
        o = orc.Oracle(cores, L, L2, paired=paired)
        o.submit(seq, q1, names, name_off, seq2, q2)
        o.finish()
        secs = time.perf_counter() - t0
        return n / secs, "port", secs, 1
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def pick_reference_threads(cores, seq, qual, names, name_off, L, n_calib=250_000, seq2=None, qual2=None, L2=0):
    """The reference's -T loop does not scale with the core count (spinlocks around parse and bucket insert), so "all the
    host threads" is not its fastest setting: time a short prefix at 1, 2, 4, ... up to the host's cores and keep the best."""
    ncpu = os.cpu_count() or 1
    cand = sorted({t for t in (1, 2, 4, 8, 16, 32, ncpu) if t <= ncpu})
    n = min(n_calib, seq.shape[0])
    res = {}
    for t in cand:
        rps, kind, _, used = run_reference_harness(cores, seq[:n], qual[:n], names[:int(name_off[n])], name_off[:n + 1], L, threads=t,
                                                   seq2=None if seq2 is None else seq2[:n], qual2=None if qual2 is None else qual2[:n], L2=L2)
        res[used] = max(rps, res.get(used, 0.0))
        if kind != "reference":   # the port has one thread only
            break
    best = max(res, key=res.get)
    return best, res


def synth_host_sample(n, L, seed, paired=False, L2=0, high_entropy=False):
    from scalce_b200 import synth
    b = synth.make_batch(n, L, seed=seed, paired=paired, L2=L2 or None, high_entropy=high_entropy)
    W = NAME_BYTES
    names = np.frombuffer(b"".join(b"SYN.%09d" % i for i in range(n)), dtype=np.uint8).copy()
    name_off = np.arange(n + 1, dtype=np.int64) * W
    return b.seq, b.qual, names, name_off, b.seq2, b.qual2


def e2e_pipelined(depth, steps_per_slot, cores, L, device, N, host_in, lib, L2=0, paired=False):
    """End-to-end throughput with `depth` steps in flight on one GPU: every slot owns a handle and a host thread and runs
    submit (H2D) -> flush -> copy-out (D2H) for its steps; the flushes take turns (one lock), the copies of different
    slots overlap (PCIe is full duplex). Every step moves all of its inputs and outputs, as in the serial arm."""
    import torch
    from scalce_b200.binding import BoostTransform
    h_seq, h_qual, h_names, h_off, h_seq2, h_qual2 = host_in
    handles = [BoostTransform(cores, L, L2, paired=paired, device=device, emit_merged=False) for _ in range(depth)]
    outbufs = [None] * depth
    sizes_seen = [None] * depth
    # one slot at a time per resource: the host->device direction (submit), the GPU (flush), the device->host direction
    # (copy-out). Without the two copy locks the slots drift into lock step - both submitting, then both copying out - and each
    # direction of the link is shared instead of the two directions being used at the same time.
    gpu_lock, h2d_lock, d2h_lock = threading.Lock(), threading.Lock(), threading.Lock()
    errors = []
    nstream = 6 if paired else 4

    def work(slot, nsteps):
        try:
            t = handles[slot]
            for _ in range(nsteps):
                t.reset_counts()
                with h2d_lock:
                    t.submit(h_seq.reshape(N, L), h_qual.reshape(N, L), h_names, h_off, h_seq2, h_qual2)
                with gpu_lock:
                    r = t.flush()
                sizes = [r.chunk_off[k][-1] for k in range(6)]
                if outbufs[slot] is None:
                    outbufs[slot] = [torch.empty(max(int(sz * 1.02), 1), dtype=torch.uint8).pin_memory() for sz in sizes]
                with d2h_lock:
                    for k in range(nstream):
                        for c in range(r.n_chunks):
                            o0, o1 = r.chunk_off[k][c], r.chunk_off[k][c + 1]
                            rc = lib.scb_copy_stream(t._h, k, c, C.c_void_p(outbufs[slot][k].data_ptr() + o0), o1 - o0)
                            if rc != 0:
                                raise RuntimeError(f"scb_copy_stream -> {rc}")
                sizes_seen[slot] = sizes
        except BaseException as ex:  # noqa: BLE001 - reported by the caller
            errors.append(ex)

    try:
        for slot in range(depth):          # warm-up: workspace and pinned buffers of every slot
            work(slot, 1)
        if errors:
            raise errors[0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(slot, steps_per_slot)) for slot in range(depth)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        total_ms = (time.perf_counter() - t0) * 1e3
        if errors:
            raise errors[0]
        # same input in every slot -> same streams
        same = all(sizes_seen[s_] == sizes_seen[0] for s_ in range(depth)) and all(
            torch.equal(outbufs[s_][1][:sizes_seen[0][1]], outbufs[0][1][:sizes_seen[0][1]]) for s_ in range(1, depth))
        if not same:
            raise RuntimeError("pipelined slots produced different streams")
        nst = depth * steps_per_slot
        return {"depth": depth, "steps": nst, "ms_per_step": total_ms / nst, "slots_agree": True}
    finally:
        for t in handles:
            t.close()


def sharded_parity_check(dist, rank, world, local, cores, L=150, n_per=150_000, by_chunk=False):
    """Run before the timed steps at N > 1: a small sharded case (n_per reads per rank, a byte budget small enough for many flush
    chunks) against the SAME concatenated input through ONE handle on rank 0's GPU (the single-GPU path, itself checked bit for
    bit against the oracle and the reference CLI fixtures by tests/). Compared: SHA-256 of every rank's slice of every stream of
    every flush chunk and of the merged streams, and of the per-read arrays. This exercises what the loopback tests cannot:
    cross-process CUDA-IPC peer stores, the side-stream row exchange and the NCCL ordering. Returns the "parity" object (rank 0).
    by_chunk: the same through scb_shard_flush (C++ orchestrator + NCCL comm library) with flush-chunk ownership, which applies
    when no merged stream is requested - the configuration of the timed steps."""
    import hashlib
    import torch
    from scalce_b200 import synth
    from scalce_b200.binding import BoostTransform
    from scalce_b200.shard import CShardedTransform, NcclCComm, ShardedTransform, TorchComm
    bsb = 16 << 20
    merged = not by_chunk

    def batch(r):
        b = synth.make_batch(n_per, L, seed=9000 + r, name_start=r * n_per)
        q = (b.qual - 33).astype(np.uint8)
        q[b.seq == ord("N")] = 0
        return b, q
    sha = lambda x: hashlib.sha256(x).hexdigest()
    b, q = batch(rank)
    t = BoostTransform(cores, L, device=local, bucket_set_bytes=bsb, emit_merged=merged)
    t.submit(b.seq, q, b.names, b.name_off)
    if by_chunk:
        os.environ["SCB_SHARD_SPLIT"] = "chunks"
        st = CShardedTransform(t, NcclCComm(dist, local), use_torch_stream=True)
    else:
        st = ShardedTransform(t, TorchComm(dist, torch.device("cuda", local)))
    try:
        res = st.flush()
    finally:
        os.environ.pop("SCB_SHARD_SPLIT", None)
    dbg = res.debug(res.n_local)
    chunks = list(range(res.n_chunks)) + ([-1] if merged else [])
    mine = {"n_chunks": res.n_chunks, "n_local": res.n_local, "rounds": st.stats["rounds"], "split": st.stats.get("split", "bucket ranges"),
            "arr": {k: sha(dbg[k].tobytes()) for k in ("node_id", "core", "end", "chunk")}, "streams": {}}
    for k in range(4):
        for c in chunks:
            x = res.stream(k, c)
            mine["streams"][f"{k}:{c}"] = (len(x), sha(x))
    t.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    out = None
    if rank == 0:
        t1 = BoostTransform(cores, L, device=local, bucket_set_bytes=bsb, emit_merged=merged)
        for r in range(world):
            bb, qq = batch(r)
            t1.submit(bb.seq, qq, bb.names, bb.name_off)
        r1 = t1.flush()
        d1 = r1.debug()
        bad = []
        if any(g["n_chunks"] != r1.n_chunks for g in gathered):
            bad.append("flush chunk count")
        o = 0
        for gi, g in enumerate(gathered):
            for k in ("node_id", "core", "end", "chunk"):
                if sha(d1[k][o:o + g["n_local"]].tobytes()) != g["arr"][k]:
                    bad.append(f"per-read {k} of rank {gi}")
            o += g["n_local"]
        for c in list(range(r1.n_chunks)) + ([-1] if merged else []):
            for k in range(4):
                whole = r1.stream(k, c)
                pos = 0
                for gi, g in enumerate(gathered):
                    ln, hs = g["streams"].get(f"{k}:{c}", (0, ""))
                    if sha(whole[pos:pos + ln]) != hs:
                        bad.append(f"stream {k} chunk {c} rank {gi}")
                    pos += ln
                if pos != len(whole):
                    bad.append(f"stream {k} chunk {c} length")
        t1.close()
        out = {"n_ranks": world, "ok": not bad, "reads": world * n_per, "read_length": L, "flush_chunks": r1.n_chunks,
               "joint_rounds": gathered[0]["rounds"], "mismatches": bad[:8], "ownership": gathered[0]["split"],
               "orchestrator": "scb_shard_flush (C++) + libscalce_b200_nccl.so" if by_chunk else "scalce_b200/shard.py + torch.distributed",
               "what": "rank-order concatenation of the ranks' streams 0-3 (every flush chunk" + (" + merged" if merged else "") + ") and per-read bucket / core / end / "
                       "chunk arrays == ONE handle fed the concatenated input, SHA-256 per rank slice; NCCL + CUDA IPC, one process per GPU"}
    return out


def bind_host_to_gpu(local):
    """N > 1: run this rank's host threads on the CPUs next to its GPU (NVML's ideal CPU affinity), before any pinned buffer
    exists, so that first-touch places the pinned pages on the GPU's NUMA node: eight ranks copying 26 GB per step through one
    socket's memory controllers and the inter-socket link is what the end-to-end arm measured in round 1 (H2D 22 GB/s and D2H
    11 GB/s per GPU instead of 55). Returns a description for the JSON line; never fails the run. SCB_BENCH_BIND=0 disables."""
    if os.environ.get("SCB_BENCH_BIND", "1") == "0":
        return {"bound": False, "why": "SCB_BENCH_BIND=0"}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        target = ideal & allowed
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:  # noqa: BLE001  (older NVML)
            pass
        if not target:
            return {"bound": False, "why": "NVML's ideal CPUs are outside this process's allowed set", "ideal_cpus": len(ideal), "allowed_cpus": len(allowed)}
        if target != allowed:
            os.sched_setaffinity(0, target)
        return {"bound": target != allowed, "cpus": len(target), "allowed_cpus": len(allowed), "numa_node": node,
                "first_cpu": min(target), "last_cpu": max(target)}
    except Exception as ex:  # noqa: BLE001
        return {"bound": False, "why": f"{type(ex).__name__}: {ex}"[:160]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configs[1..4]; the default c2 is the one the metric is quoted on")
    ap.add_argument("--cores", type=int, default=2048, help="2048 = the shared-memory resident headline core set; >= 100000 = that many cores of 10-14 bases")
    ap.add_argument("--reads", type=int, default=0, help="reads (pairs) per GPU per step; default: the config's")
    ap.add_argument("--length", type=int, default=0)
    ap.add_argument("--flushes", type=int, default=0, help="flushes per step (the job's populations carry over); default: 1, or as many as the GPU's memory asks for")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads of the CPU baseline sample (default: ~2M at 150 bp, scaled by read length and core set)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-depth", type=int, default=2,
                    help="1 GPU only: steps in flight in the end-to-end arm. 1 = one step after the other; 2 (default) = two handles on two host "
                         "threads, so step k+1's H2D runs while step k's result drains over the other PCIe direction (flushes serialised)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded parity check that precedes the timed steps")
    ap.add_argument("--orchestrator", default="cpp", choices=["cpp", "python"],
                    help="N > 1: who runs the sharded sequence - the library (scb_shard_flush over libscalce_b200_nccl.so, C++ + NCCL, the default) "
                         "or scalce_b200/shard.py over torch.distributed")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cf = CONFIGS[a.config]
    L = a.length or cf["L"]
    paired, L2, high_entropy = cf["paired"], cf["L2"], cf["high_entropy"]
    if paired and a.length:
        L2 = L
    N = a.reads or (cf["reads"] if "reads" in cf else cf["total"] // max(world, 1))
    cores, core_desc, mean_core = bench_cores(a.cores)
    workload = cf["desc"].format(n=(f"{N / 1e6:g}")) + ", " + core_desc
    bpr = algorithmic_bytes_per_read(L, NAME_BYTES, mean_core, paired, L2)
    cpu_sample = a.cpu_sample or max(100_000, int(2_000_000 * 150 / (L + (L2 if paired else 0)) / (8 if a.cores >= 100000 else 1)))

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        ns = min(cpu_sample, N)
        seq, qual, names, name_off, seq2, qual2 = synth_host_sample(ns, L, seed=1, paired=paired, L2=L2, high_entropy=high_entropy)
        kw = dict(seq2=seq2, qual2=qual2, L2=L2 if paired else 0)
        T, calib = pick_reference_threads(cores, seq, qual, names, name_off, L, **kw)
        vals = []
        for s in range(a.warmup + a.steps):
            rps, kind, secs, used = run_reference_harness(cores, seq, qual, names, name_off, L, threads=T, **kw)
            if s >= a.warmup:
                vals.append((rps, secs))
            if s == 0 and secs * (a.warmup + a.steps) > 240:  # keep the run within a few minutes
                vals = [(rps, secs)]
                break
        rps = float(np.mean([v[0] for v in vals]))
        ms = float(np.mean([v[1] for v in vals])) * 1e3
        sample = (f"first {ns} reads of the workload per step, in-memory, {used} thread(s) = the fastest of {sorted(calib)} on this host "
                  f"({os.cpu_count()} cores; the reference serialises parse and bucket insert under spinlocks; bit-exact only at 1 thread)")
        print(json.dumps({
            "impl": "reference", "metric": "reads/s of core-scan+bucket+reorder", "value": rps, "unit": "reads/s", "n_gpus": a.gpus,
            "steps": len(vals), "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cf["scaling"], "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "bases_per_s": rps * (L + (L2 if paired else 0)),
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": rps, "unit": "reads/s", "cores": used, "kind": kind, "sample": sample,
                             "reads_per_s_by_threads": {str(k): v for k, v in sorted(calib.items())}},
            "e2e": {"value": rps, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ native arm (B200)
    import torch
    from scalce_b200 import synth
    from scalce_b200.binding import BoostTransform, load_library

    lib = load_library()  # fails loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    # N = 1 keeps the whole host (the CPU baseline leg uses every core, and one GPU's copies already run at the PCIe rate)
    host_binding = bind_host_to_gpu(local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- N > 1: parity of the sharded path on this hardware, before anything is timed ----------------------------
    parity = None
    if world > 1 and not a.no_parity:
        parity = sharded_parity_check(dist, rank, world, local, cores)
        torch.cuda.synchronize()
        dist.barrier()
        if a.orchestrator == "cpp":
            p2 = sharded_parity_check(dist, rank, world, local, cores, by_chunk=True)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                parity = {"ok": bool(parity["ok"] and p2["ok"]), "bucket_ranges": parity, "flush_chunks": p2}

    d = synth.make_batch_cuda(N, L, seed=1 + rank, paired=paired, L2=L2 or None, high_entropy=high_entropy, device=f"cuda:{local}")
    seq, qual, names, name_off = d["seq"], d["qual"], d["names"], d["name_off"]
    # host side of the boundary prepares quality payload (output_quality: q - 33, 0 under N)
    qual = torch.where(seq == ord("N"), torch.zeros_like(qual), qual - 33)
    seq2 = qual2 = None
    if paired:
        seq2 = d["seq2"]
        qual2 = torch.where(seq2 == ord("N"), torch.zeros_like(d["qual2"]), d["qual2"] - 33)
    del d
    torch.cuda.synchronize()
    torch.cuda.empty_cache()          # the generator's temporaries go back to the driver: the library allocates with cudaMalloc, not through torch

    tr = BoostTransform(cores, L, L2, paired=paired, device=local, emit_merged=False)
    info = tr.table_info()
    engine = {0: "dense (shared-memory population rows)", 1: "sparse (bucket-major candidate lists)", 2: "sequential"}[tr.resolve_engine]
    sharded = None
    if world > 1:
        # the global input is the concatenation of the ranks' batches in rank order; results are the single-GPU
        # (= reference -T 1) order of that input, each rank emitting a contiguous slice of the bucket order
        from scalce_b200.shard import CShardedTransform, NcclCComm, ShardedTransform, TorchComm
        if a.orchestrator == "cpp":
            sharded = CShardedTransform(tr, NcclCComm(dist, local), use_torch_stream=True)   # library work + NCCL on torch's current stream: the events below see it
        else:
            sharded = ShardedTransform(tr, TorchComm(dist, dev))

    # flushes per step: one, unless the flush workspace would not fit next to the resident inputs
    F = a.flushes
    if F <= 0:
        F = 1
        free_b, _ = torch.cuda.mem_get_info(dev)
        per_read = 360 + 3 * L + (L // 4) + ((3 * L2) if paired else 0)     # workspace + output streams + the concatenation copy of a multi-batch flush, generous
        while F < 16 and (N / F) * per_read > 0.85 * free_b:
            F += 1
        if world > 1:
            F = 1
    edges = [N * k // F for k in range(F + 1)]

    # name offsets of a slice must start at 0: fixed-width names, so one shared offsets array serves every slice
    off0 = name_off

    def one_step_sharded():
        hr = time.perf_counter()
        tr.reset_counts()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        tr.submit_device(N, seq.data_ptr(), qual.data_ptr(), names.data_ptr(), name_off.data_ptr(),
                         seq2.data_ptr() if paired else None, qual2.data_ptr() if paired else None)
        h1 = time.perf_counter()
        sharded.flush()
        h2 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        h3 = time.perf_counter()
        st = dict(sharded.stats["ms"])
        st["wall:_reset"] = (h0 - hr) * 1e3
        st["wall:_sync_after"] = (h3 - h2) * 1e3
        st["wall:_submit"] = (h1 - h0) * 1e3
        st["wall:_flush_call"] = (h2 - h1) * 1e3
        st["_rounds"] = sharded.stats["rounds"]
        st["_split"] = sharded.stats.get("split", "bucket ranges")
        for k, v in (sharded.stats.get("wall_ms") or {}).items():
            st["wall:" + k] = v
        return e0.elapsed_time(e1), st

    def one_step():
        if sharded is not None:
            return one_step_sharded()
        # one compression job per step: same handle (automaton + device workspace), populations reset
        tr.reset_counts()
        ms, st = 0.0, {}
        for k in range(F):
            a0, a1 = edges[k], edges[k + 1]
            tr.submit_device(a1 - a0, seq[a0:a1].data_ptr(), qual[a0:a1].data_ptr(), names[a0 * NAME_BYTES:].data_ptr(), off0.data_ptr(),
                             seq2[a0:a1].data_ptr() if paired else None, qual2[a0:a1].data_ptr() if paired else None)
            # several flushes per step (a job beyond one GPU's memory): all but the last are streaming flushes - complete flush chunks
            # only, the open chunk's reads stay pending - so the step's chunks are those of ONE flush over the whole input
            r = tr.flush() if k == F - 1 else tr.flush_closed()
            ms += r.device_ms
            for kk, v in tr.stage_ms().items():
                st[kk] = st.get(kk, 0.0) + v
        st["_rounds"] = tr.resolve_rounds
        return ms, st

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        one_step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = lib.scb_kernel_launches(None)
    t0 = time.perf_counter()
    dev_ms, stages = [], []
    for _ in range(a.steps):
        ms, st = one_step()
        dev_ms.append(ms); stages.append(st)
        if os.environ.get("SCB_BENCH_VERBOSE"):
            print("step", ms, st, file=sys.stderr)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    launches = lib.scb_kernel_launches(None) - launches0
    clocks = sampler.stop() if sampler else None
    ms_step = float(np.mean(dev_ms))
    by_rank = None
    if dist is not None:
        # every rank's own view of the step: device time, host wall per phase of scb_shard_flush (collectives and waiting included)
        mine = {"rank": rank, "ms_per_step": ms_step,
                "wall_ms": {k[5:]: round(float(np.mean([s_[k] for s_ in stages])), 3) for k in stages[0] if k.startswith("wall:")}}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        by_rank = gathered
        tt = torch.tensor([ms_step, wall_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step, wall_ms = float(tt[0]), float(tt[1])
    value = N * world / (ms_step * 1e-3)
    workspace_bytes = tr.device_bytes

    # ---- end to end through the C ABI with host buffers -------------------------------------
    e2e = None
    h2d = int(sum(x.numel() * x.element_size() for x in (seq, qual, names, name_off) + ((seq2, qual2) if paired else ())))
    if not a.no_e2e and h2d > (48 << 30):
        e2e = {"value": None, "unit": "reads/s", "skipped": f"{h2d / 2**30:.0f} GiB of pinned host input per step: beyond what this arm pins"}
    elif not a.no_e2e:
        try:
            dev_in = [seq, qual, names, name_off] + ([seq2, qual2] if paired else [])
            hs = [x.cpu().pin_memory() for x in dev_in]
            hn = [x.numpy() for x in hs]
            h_seq, h_qual, h_names, h_off = hn[:4]
            h_seq2, h_qual2 = (hn[4].reshape(N, L2), hn[5].reshape(N, L2)) if paired else (None, None)
            nstream = 6 if paired else 4
            outbuf = None
            e_steps = max(1, a.e2e_steps)
            e_ms, e_parts = [], []
            # ONE handle (and, at N > 1, one ShardedTransform: receive arrays, IPC mappings, workspace) for all steps:
            # nothing is allocated or mapped inside the timed steps after the warm-up step
            t, e_sh = tr, sharded
            for s in range(1 + e_steps):
                barrier()
                t1 = time.perf_counter()
                t.reset_counts()
                t.submit(h_seq.reshape(N, L), h_qual.reshape(N, L), h_names, h_off, h_seq2, h_qual2)
                t2 = time.perf_counter()
                r = e_sh.flush() if e_sh is not None else t.flush()
                t3 = time.perf_counter()
                sizes = [r.chunk_off[k][-1] for k in range(6)]
                if outbuf is None or any(outbuf[k].numel() < sizes[k] for k in range(6)):
                    outbuf = [torch.empty(max(int(sz * 1.02), 1), dtype=torch.uint8).pin_memory() for sz in sizes]
                for k in range(nstream):
                    for c in range(r.n_chunks):
                        o0, o1 = r.chunk_off[k][c], r.chunk_off[k][c + 1]
                        lib.scb_copy_stream(t._h, k, c, C.c_void_p(outbuf[k].data_ptr() + o0), o1 - o0)
                torch.cuda.synchronize()
                t4 = time.perf_counter()
                if s >= 1:
                    e_ms.append((t4 - t1) * 1e3)
                    e_parts.append(((t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
            e_med, e_min = float(np.median(e_ms)), float(np.min(e_ms))
            if dist is not None:
                tt = torch.tensor([e_med, e_min], device=f"cuda:{local}", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e_med, e_min = float(tt[0]), float(tt[1])
            d2h = int(sum(sizes))
            e2e = {"value": N * world / (e_med * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e_med, "ms_per_step_min": e_min, "steps": e_steps, "statistic": "median over the steps (max over ranks)",
                   "ms_submit_flush_copyout": [float(np.median([p[i] for p in e_parts])) for i in range(3)],
                   "note": "pinned host buffers -> scb_submit (H2D) -> scb_flush -> scb_copy_stream of every stream (D2H); one handle" +
                           (" and one sharded transform (receive arrays, IPC mappings)" if world > 1 else "") + " reused across steps"}
            if host_binding is not None:
                e2e["host_binding_rank0"] = host_binding
            tr.close()
            tr = None
            if a.e2e_depth > 1 and world == 1:
                del outbuf
                try:
                    piped = e2e_pipelined(a.e2e_depth, max(2, (e_steps + 1) // 2), cores, L, local, N, (h_seq, h_qual, h_names, h_off, h_seq2, h_qual2), lib, L2, paired)
                except Exception as ex:  # the serial figure stands
                    piped = {"error": repr(ex)}
                e2e["serial"] = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "ms_per_step_min": e_min}
                if "ms_per_step" in piped and piped["ms_per_step"] < e2e["ms_per_step"]:
                    e2e["value"] = N / (piped["ms_per_step"] * 1e-3)
                    e2e["ms_per_step"] = piped["ms_per_step"]
                    e2e["steps"] = piped["steps"]
                    e2e["statistic"] = "wall time of all steps / steps"
                    e2e["note"] += f"; {a.e2e_depth} steps in flight (one handle + host thread each, flushes serialised): every step still copies its inputs in and its streams out"
                e2e["pipelined"] = piped
            del hs, hn
        except Exception as ex:   # the kernel-only line above stands; say what happened instead of dying
            if world > 1:
                raise           # a rank that drops out of the collectives would hang the others
            e2e = {"value": None, "unit": "reads/s", "error": repr(ex)}
    if tr is not None:
        tr.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage + whole-path figure --------------------------------------
    peak, peak_src = measured_peak_gbs()
    split_mode = stages[-1].get("_split") if isinstance(stages[-1], dict) else None
    mean_st = {k: float(np.mean([s[k] for s in stages])) for k in stages[0] if k != "_split"}
    resolve_rounds = mean_st.pop("_rounds", None)
    # N > 1: host wall time per phase of scb_shard_flush on rank 0 (device work + collectives + waiting for the other ranks)
    phase_wall = {k[5:]: mean_st.pop(k) for k in list(mean_st) if k.startswith("wall:")} or None
    dom = max(mean_st, key=mean_st.get)
    packed = (L - mean_core + 3) // 4 + (2 if L > 255 else 1)
    PWB = (L + 15) // 16 * 4                       # bytes of a 2-bit packed row
    pair_b = (3 * L2 + (L2 + 3) // 4) if paired else 0
    stage_bytes = {  # algorithmic bytes per read of each stage
        "scan": L + 8, "resolve": 16, "chunks": 8, "sort": 24, "ties": 0,
        # emit reads: quality row L + 2-bit row + name + metadata word; writes: quality row L + packed record + name record (+ mate 2)
        "emit": (L + PWB + NAME_BYTES + 8) + (L + packed + NAME_BYTES + 1) + pair_b, "merged": 0, "arrays": 16,
        # sharded run: pack + exchange move the payload once each (aux word, packed row, quality row, name)
        "finalize": 8, "hist": 4, "resolve_rounds": 16, "pack": 2 * (8 + PWB + L + NAME_BYTES), "exchange": 2 * (8 + PWB + NAME_BYTES),
        "exchange_rows": 2 * L + (4 * L2 if paired else 0), "import": 16,
        "emit_early": 2 * (NAME_BYTES + 1) + packed + PWB + 8,
    }
    ach = N * stage_bytes.get(dom, 0) / (mean_st[dom] * 1e-3) / 1e9
    # DRAM bytes per read of each stage's kernels (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full pass over the
    # headline step, profiles/r02_launches_summary.txt); only valid for the headline shape and core set
    ncu_traffic_per_read = NCU_TRAFFIC_PER_READ
    traffic = float(ncu_traffic_per_read[dom]) * N if (a.config == "c2" and a.cores == 2048 and L == 150 and world == 1 and dom in ncu_traffic_per_read) else None
    per_stage = {k: {"ms": v, "bytes_per_read": stage_bytes.get(k), "achieved_gbs": (N * stage_bytes[k] / (v * 1e-3) / 1e9) if (v > 0 and stage_bytes.get(k)) else None}
                 for k, v in mean_st.items()}
    for k, d_ in per_stage.items():
        d_["frac"] = (d_["achieved_gbs"] / peak) if d_["achieved_gbs"] else None
    roof = {"bound": "hbm", "kernel": dom, "kernels": STAGE_KERNELS.get(dom), "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": NCU_TRAFFIC_SOURCE if traffic else None,
            "peak_source": peak_src, "bytes_per_read": stage_bytes.get(dom), "stage_ms": mean_st, "per_stage": per_stage, "resolve_rounds": resolve_rounds,
            **({"phase_wall_ms_rank0": phase_wall} if phase_wall else {}), **({"by_rank": by_rank} if by_rank else {})}
    pipe = N * bpr / (ms_step * 1e-3) / 1e9
    pipeline = {"achieved": pipe, "unit": "GB/s", "frac_of_peak": pipe / peak, "frac_of_nominal_8TBs": pipe / 8000.0, "bytes_per_read": bpr}

    cpu = None
    if not a.no_cpu:
      try:
        ns = min(cpu_sample, N)
        s_seq = seq[:ns].cpu().numpy(); s_qual = (qual[:ns] + 33).cpu().numpy()
        s_names = names[:ns * NAME_BYTES].cpu().numpy(); s_off = name_off[:ns + 1].cpu().numpy()
        kw = dict(seq2=seq2[:ns].cpu().numpy(), qual2=(qual2[:ns] + 33).cpu().numpy(), L2=L2) if paired else {}
        T, calib = pick_reference_threads(cores, s_seq, s_qual, s_names, s_off, L, **kw)
        rps1, kind, secs1, _ = run_reference_harness(cores, s_seq, s_qual, s_names, s_off, L, threads=1, **kw)
        rps, secs, used = rps1, secs1, 1
        if T > 1:
            rps, kind, secs, used = run_reference_harness(cores, s_seq, s_qual, s_names, s_off, L, threads=T, **kw)
            if rps < rps1:
                rps, secs, used = rps1, secs1, 1
        cpu = {"value": rps, "unit": "reads/s", "cores": used, "kind": kind,
               "sample": f"first {ns} reads of rank 0's batch, in-memory, {used} thread(s) = the reference's fastest setting on this host ({secs:.1f} s); "
                         f"1 thread (its only bit-exact mode): {rps1:.0f} reads/s ({secs1:.1f} s)",
               "value_1thread": rps1, "reads_per_s_by_threads": {str(k): v for k, v in sorted(calib.items())}, "host_cores_available": os.cpu_count()}
      except Exception as ex:   # the GPU line stands
        cpu = {"value": None, "unit": "reads/s", "error": repr(ex)}

    out = {
        "metric": "reads/s of core-scan+bucket+reorder", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": cf["scaling"],
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "bases_per_s": value * (L + (L2 if paired else 0)),
        "config": {"workload": workload, "name": a.config, "reads_per_gpu": N, "read_length": L, "paired": paired, "mate2_length": L2 if paired else 0,
                   "high_entropy_qualities": high_entropy, "flushes_per_step": F,
                   "core_set": {"cores": len(cores), "automaton_states": info["n_states"], "buckets": info["n_buckets"],
                                "table": "shared memory (u16 transitions)" if info["smem_resident"] else "global memory / L2 (u32 transitions)", "tie_break_engine": engine},
                   "l2": f"inputs ({h2d / 1e9:.1f} GB per step) far exceed the 126 MB L2",
                   "timing": ("CUDA events on the library stream around scb_flush; one handle, bucket populations reset between steps" if world == 1 else
                              "CUDA events on the rank's stream (library work and NCCL collectives are ordered on it; the side stream of the row exchange is joined before the emit) around submit + sharded flush, max over ranks"),
                   "orchestrator": ("n/a" if world == 1 else ("scb_shard_flush (C++) + libscalce_b200_nccl.so" if a.orchestrator == "cpp" else "scalce_b200/shard.py + torch.distributed")),
                   "multi_gpu": ("n/a" if world == 1 else "contiguous input shards; joint exact tie-break (all-gather of bucket histograms per round), " +
                                 ("whole flush chunks handed to the ranks next to them: only the reads of chunks at shard edges move" if split_mode == "flush chunks"
                                  else "bucket-range exchange of packed reads + qualities + names") +
                                 " by peer stores over NVLink (CUDA IPC); output = single-GPU order of the concatenated input"),
                   "ownership": None if world == 1 else (split_mode or "bucket ranges")},
        "roofline": roof, "pipeline_roofline": pipeline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "workspace_bytes_per_flush": int(workspace_bytes),
    }
    if parity is not None:
        out["parity"] = parity
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

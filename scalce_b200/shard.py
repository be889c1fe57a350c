"""Sharded (multi-GPU) run of the boosting transform: one process per GPU, reads sharded as
contiguous input slices, results bit-identical to one GPU fed the whole input.

The library (include/scalce_b200.h, "Sharded run") does all per-GPU work; this module is the
plumbing between ranks:
  * flush-chunk numbering: a (carry, chunk) pair handed from rank to rank,
  * tie-break: rank 0 resolves its shard alone, then all later shards iterate ONE global fixed
    point together - per round each rank re-decides its reads from the populations of everything
    before it under the current global assignment and the ranks all-gather their bucket histograms
    (n_buckets+2 words each); the run stops at the first round (other than the first) in which no
    read on any rank changed, which proves the assignment is the sequential one,
  * exchange: all-reduce of the bucket histogram, split of the bucket emission order into
    contiguous slices balanced by reads, one all-to-all per payload array over NCCL,
  * each rank then sorts and emits its slice: rank-order concatenation of the ranks' streams is,
    per flush chunk, the single-GPU (= reference -T 1) stream.

Comm back-ends: TorchComm (torch.distributed: NCCL on GPUs, gloo for the CPU tests of the host
logic) and LoopbackComm (ranks are threads of one process sharing one GPU; used to test the
whole sharded path on a single-GPU box).
"""
from __future__ import annotations

import ctypes as C
import sys
import threading

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous, balanced slices: the first n % world ranks get one extra read."""
    base, extra = divmod(n, world)
    b = [0]
    for g in range(world):
        b.append(b[-1] + base + (1 if g < extra else 0))
    return b


def max_over_ranks(values, dist=None, device=None):
    """Element-wise MAX of a list of floats over all ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def balanced_split(hist, world: int):
    """Splits the bucket emission order [0, len(hist)) into `world` contiguous slices with about equal
    read counts. Returns split[world+1], split[0] = 0, split[world] = len(hist); rank g owns
    [split[g], split[g+1]). A bucket is never divided (its reads must be sorted together)."""
    hist = np.asarray(hist, dtype=np.int64)
    ncols = int(hist.shape[0])
    total = int(hist.sum())
    cum = np.cumsum(hist)
    split = [0]
    for g in range(1, world):
        target = (total * g) // world
        # first bucket whose inclusive cumulative count exceeds the target starts rank g's slice ... unless the
        # previous slice would then be empty while this bucket alone overshoots: keep boundaries monotone
        k = int(np.searchsorted(cum, target, side="right"))
        k = min(max(k, split[-1]), ncols)
        split.append(k)
    split.append(ncols)
    return split


def chain_chunks(sizes_fn, comm):
    """Flush-chunk numbering along the global order: ranks take turns, each passing (carry, chunk) on.
    sizes_fn(carry_in, chunk_in) -> (carry_out, chunk_out) is this rank's scb_shard_sizes.
    Returns the global number of flush chunks."""
    carry, chunk = 0, 0
    for g in range(comm.world):
        if g == comm.rank:
            carry, chunk = sizes_fn(carry, chunk)
        carry, chunk = comm.bcast_host([int(carry), int(chunk)], src=g)
    return chunk + (1 if carry > 0 else 0)


# ---------------------------------------------------------------------------------------------
# communication back-ends
# ---------------------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed plumbing (NCCL for device tensors; works on gloo/CPU for the host-logic tests)."""

    def __init__(self, dist, device):
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def allgather(self, t):
        import torch
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat)
        return out.view((self.world,) + tuple(t.shape))

    def allgather_host(self, ints):
        import torch
        t = torch.tensor(list(ints), dtype=torch.int64, device=self.device)
        return self.allgather(t).cpu().tolist()

    def bcast_host(self, ints, src):
        import torch
        t = torch.tensor(list(ints), dtype=torch.int64, device=self.device)
        self.dist.broadcast(t, src=src)
        return t.cpu().tolist()

    def allgather_bytes(self, b: bytes):
        import torch
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)
        return [bytes(x.cpu().numpy().tobytes()) for x in self.allgather(t)]

    same_process = False

    def all_to_all_bytes(self, send, send_counts, recv_counts, slack=0):
        """send: uint8 tensor laid out destination-major; counts in bytes. Returns the receive tensor
        (source-major) with `slack` extra bytes allocated past the end."""
        import torch
        total = int(sum(recv_counts))
        buf = torch.empty(total + slack, dtype=torch.uint8, device=send.device)
        self.dist.all_to_all_single(buf[:total], send[: int(sum(send_counts))], [int(x) for x in recv_counts], [int(x) for x in send_counts])
        return buf

    def map_peers(self, transform, ptrs, changed):
        """Publishes this rank's receive arrays (device pointers, 0 = unused) and returns table[g][k]: rank g's
        array k as mapped into THIS process (CUDA IPC over NVLink peer access; own rank: the local pointer).
        Handles are exchanged only when some rank's arrays moved; mappings are cached."""
        import torch
        from .binding import _check, load_library
        L = load_library()
        if not hasattr(self, "_ipc"):
            self._ipc = {}        # (g, k) -> (handle bytes, mapped pointer)
            self._table = None
        any_changed = max(x[0] for x in self.allgather_host([1 if (changed or self._table is None) else 0]))
        if any_changed:
            nk = len(ptrs)
            hb = np.zeros((nk, 65), dtype=np.uint8)
            for k, p in enumerate(ptrs):
                if p:
                    buf = (C.c_uint8 * 64)()
                    _check(L.scb_ipc_export(transform._h, C.c_void_p(p), buf))
                    hb[k, :64] = np.frombuffer(buf, dtype=np.uint8)
                    hb[k, 64] = 1
            allh = self.allgather(torch.from_numpy(hb).to(self.device)).cpu().numpy()
            table = []
            for g in range(self.world):
                row = []
                for k in range(nk):
                    if g == self.rank:
                        row.append(int(ptrs[k] or 0))
                        continue
                    if not allh[g, k, 64]:
                        row.append(0)
                        continue
                    hbytes = allh[g, k, :64].tobytes()
                    old = self._ipc.get((g, k))
                    if old is None or old[0] != hbytes:
                        if old is not None:
                            _check(L.scb_ipc_close(transform._h, C.c_void_p(old[1])))
                        out = C.c_void_p()
                        hbuf = (C.c_uint8 * 64).from_buffer_copy(hbytes)
                        _check(L.scb_ipc_open(transform._h, hbuf, C.byref(out)))
                        self._ipc[(g, k)] = (hbytes, out.value)
                    row.append(self._ipc[(g, k)][1])
                table.append(row)
            self._table = table
        else:
            self._table[self.rank] = [int(p or 0) for p in ptrs]
        return self._table

    def barrier(self):
        self.dist.barrier()


class _LoopbackShared:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class LoopbackComm:
    """`world` ranks as threads of ONE process (one GPU): collectives are shared-memory hand-offs and
    device-to-device copies. Same interface as TorchComm, so the orchestration code is the same."""

    def __init__(self, shared: _LoopbackShared, rank: int, device):
        self.s = shared
        self.rank = rank
        self.world = shared.world
        self.device = device

    @staticmethod
    def make(world, device):
        sh = _LoopbackShared(world)
        return [LoopbackComm(sh, r, device) for r in range(world)]

    def _exchange(self, obj):
        self.s.slots[self.rank] = obj
        self.s.barrier.wait()
        got = list(self.s.slots)
        self.s.barrier.wait()
        return got

    def allgather(self, t):
        import torch
        torch.cuda.synchronize() if t.is_cuda else None
        return torch.stack(self._exchange(t.clone()))

    def allgather_host(self, ints):
        return [list(x) for x in self._exchange(list(ints))]

    def bcast_host(self, ints, src):
        return list(self._exchange(list(ints))[src])

    def allgather_bytes(self, b: bytes):
        return [bytes(x) for x in self._exchange(bytes(b))]

    same_process = True

    def all_to_all_bytes(self, send, send_counts, recv_counts, slack=0):
        import torch
        if send.is_cuda:
            torch.cuda.synchronize()
        self.s.slots[self.rank] = (send, [int(x) for x in send_counts])
        self.s.barrier.wait()
        total = int(sum(recv_counts))
        buf = torch.empty(total + slack, dtype=torch.uint8, device=send.device)
        o = 0
        for src in range(self.world):
            t, cnts = self.s.slots[src]
            a = int(sum(cnts[: self.rank]))
            n = cnts[self.rank]
            assert n == int(recv_counts[src])
            buf[o:o + n] = t[a:a + n]
            o += n
        if send.is_cuda:
            torch.cuda.synchronize()
        self.s.barrier.wait()   # senders may release their buffers only after every receiver copied
        return buf

    def map_peers(self, transform, ptrs, changed):
        return [[int(p or 0) for p in row] for row in self._exchange(list(ptrs))]   # one address space

    def barrier(self):
        self.s.barrier.wait()


# ---------------------------------------------------------------------------------------------
# the sharded transform
# ---------------------------------------------------------------------------------------------
class _DevPtr:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _dev_bytes(ptr, nbytes, device):
    import torch
    if nbytes <= 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


class ShardedTransform:
    """One rank of a sharded run. `transform` is this rank's BoostTransform (its device = this rank's GPU);
    submit this rank's slice of the input to it as usual, then call flush() on every rank."""

    def __init__(self, transform, comm, use_torch_stream=True, p2p=True, overlap=True):
        """p2p=True: fused pack + send over peer memory (CUDA IPC / NVLink stores from the gather kernels);
        p2p=False: rows are staged locally and moved by the comm's all-to-all (NCCL)."""
        self.t = transform
        self.comm = comm
        self.p2p = p2p and hasattr(comm, "map_peers")
        self.on_torch_stream = bool(use_torch_stream)
        self.overlap = overlap      # row exchange overlapped with the receive side's sort (p2p only)
        self.stats = {}
        self._keep = None
        if use_torch_stream:
            import torch
            from .binding import _check, load_library
            with torch.cuda.device(transform.cfg.device):
                s = torch.cuda.current_stream().cuda_stream
            _check(load_library().scb_set_stream(transform._h, C.c_void_p(s), 1))

    def flush(self):
        import torch
        from .binding import FlushResult, ScbResult, ScbShardPeer, ScbShardXfer, _check, load_library
        L = load_library()
        h = self.t._h
        cfg = self.t.cfg
        comm = self.comm
        G, r = comm.world, comm.rank
        dev = torch.device("cuda", cfg.device)
        L1, L2 = cfg.read_length[0], cfg.read_length[1]
        ms = {}

        def lap(name):
            ms[name] = ms.get(name, 0.0) + float(L.scb_shard_last_ms(h))

        ncols_c, rootpos_c = C.c_int32(), C.c_int32()
        _check(L.scb_shard_info(h, C.byref(ncols_c), C.byref(rootpos_c)))
        ncols = ncols_c.value

        # ---- scan -------------------------------------------------------------------------------------
        n_local = C.c_int64()
        _check(L.scb_shard_scan(h, C.byref(n_local)))
        lap("scan")
        ns = [x[0] for x in comm.allgather_host([n_local.value])]
        before = [0]
        for x in ns:
            before.append(before[-1] + x)
        n_global = before[-1]

        # ---- flush chunks along the global order ------------------------------------------------------------
        def sizes(carry, chunk):
            co, ko = C.c_uint64(), C.c_int32()
            _check(L.scb_shard_sizes(h, carry, chunk, C.byref(co), C.byref(ko)))
            lap("chunks")
            return co.value, ko.value
        n_chunks = chain_chunks(sizes, comm)

        # ---- tie-break -------------------------------------------------------------------------------------
        tot = torch.zeros(ncols + 1, dtype=torch.int32, device=dev)   # u32 bit patterns
        sync = (lambda: None) if self.on_torch_stream else (lambda: torch.cuda.synchronize(dev))   # one stream orders everything
        if r == 0:
            _check(L.scb_shard_resolve_local(h, C.c_void_p(tot.data_ptr())))
            lap("resolve")
        rounds = 0
        sync()
        allt = comm.allgather(tot)
        if G > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            # Rounds are enqueued one ahead of the convergence check: the "changed" count of round k is read back
            # (pinned buffer + event) while round k+1 is already running, so the device never waits for the host. A round
            # launched after the fixed point re-decides everything from the same inputs and changes nothing.
            pipelined = self.on_torch_stream and not isinstance(comm, LoopbackComm)
            inflight = []          # (event, pinned changed count, was_first)
            if pipelined and not hasattr(self, "_pin"):
                self._pin = [torch.empty(1, dtype=torch.int64, pin_memory=True) for _ in range(4)]   # allocated once: pinning is slow
            first = True
            done = False
            while not done:
                if r > 0:
                    if first:   # guess: rank 0's histogram scaled to the reads before this shard (exact for rank 1)
                        if ns[0] > 0:
                            bf = (allt[0, :ncols].to(torch.int64) * before[r] // ns[0]).to(torch.int32)
                        else:
                            bf = torch.zeros(ncols, dtype=torch.int32, device=dev)
                    else:
                        bf = allt[:r, :ncols].sum(0, dtype=torch.int32)    # u32 bit patterns: wrap-around sum is the u32 sum
                    sync()
                    _check(L.scb_shard_resolve_round(h, C.c_void_p(bf.data_ptr()), before[r], 1 if first else 0, C.c_void_p(tot.data_ptr())))
                    sync()
                allt = comm.allgather(tot)
                rounds += 1
                chg = allt[1:, ncols].sum(dtype=torch.int64)
                if pipelined:
                    hb = self._pin[rounds % 4]
                    hb.copy_(chg.view(1), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    inflight.append((ev, hb, first, bf if r > 0 else None))   # keeps bf alive until its round has run
                    if len(inflight) == 2:
                        ev0, hb0, f0, _ = inflight.pop(0)
                        ev0.synchronize()
                        done = (not f0) and int(hb0[0]) == 0
                else:
                    done = (not first) and int(chg.item()) == 0
                first = False
            e1.record()
            torch.cuda.synchronize(dev)
            ms["resolve_rounds"] = e0.elapsed_time(e1)
        gtot = allt[:, :ncols].sum(0, dtype=torch.int64).to(torch.int32).contiguous()
        torch.cuda.synchronize(dev)
        _check(L.scb_shard_finalize(h, C.c_void_p(gtot.data_ptr()), n_global))
        lap("finalize")

        # ---- bucket-range split ----------------------------------------------------------------------------
        hist = torch.zeros(ncols, dtype=torch.int32, device=dev)
        _check(L.scb_shard_bucket_hist(h, C.c_void_p(hist.data_ptr())))
        lap("hist")
        torch.cuda.synchronize(dev)
        ghist = comm.allgather(hist).to(torch.int64).sum(0).cpu().numpy()
        split = balanced_split(ghist, G)
        split_c = (C.c_int64 * (G + 1))(*split)
        x = ScbShardXfer()
        if self.p2p:
            _check(L.scb_shard_partition(h, split_c, G, C.byref(x)))
        else:
            _check(L.scb_shard_pack(h, split_c, G, C.byref(x)))
        lap("pack")

        # ---- exchange ---------------------------------------------------------------------------------------
        cr = [int(x.cnt_reads[g]) for g in range(G)]
        cn = [int(x.cnt_name_bytes[g]) for g in range(G)]
        mat = comm.allgather_host(cr + cn)
        rr = [mat[s][r] for s in range(G)]
        rn = [mat[s][G + r] for s in range(G)]
        n_recv, nb_recv = sum(rr), sum(rn)
        prow = x.packed_row_bytes
        y = ScbShardXfer()
        y.n, y.name_bytes, y.packed_row_bytes = n_recv, nb_recv, prow
        keep = {}
        if self.p2p:
            # fused pack + send: the row gathers write straight into the owners' receive arrays over NVLink
            need = [n_recv * 8, n_recv * prow + 64, n_recv * L1 if cfg.use_quals else 0, nb_recv + 16 if cfg.use_names else 0,
                    n_recv * L2 if cfg.paired else 0, n_recv * L2 if (cfg.paired and cfg.use_quals) else 0]
            ptrs = (C.c_void_p * 6)()
            chg = C.c_int32()
            _check(L.scb_shard_recv_reserve(h, (C.c_int64 * 6)(*need), ptrs, C.byref(chg)))
            table = comm.map_peers(self.t, [ptrs[k] for k in range(6)], bool(chg.value))
            peers = (ScbShardPeer * G)()
            for g in range(G):
                pg = peers[g]
                pg.aux, pg.packed, pg.qual1, pg.names, pg.seq2, pg.qual2 = [table[g][k] or None for k in range(6)]
                pg.row_off = sum(mat[s][g] for s in range(r))
                pg.name_off = sum(mat[s][G + g] for s in range(r))
            y.aux, y.packed, y.qual1, y.names, y.seq2, y.qual2 = [ptrs[k] for k in range(6)]
            if self.overlap:
                # what the receive side needs to SORT goes first; the quality / mate-2 rows then cross NVLink on a
                # side stream while the received reads are sorted, and are only awaited before the emit
                _check(L.scb_shard_send(h, r, G, peers, 1, 0))
                lap("exchange")
                comm.barrier()
                _check(L.scb_shard_import(h, C.byref(y), n_chunks))
                lap("import")
                _check(L.scb_shard_send(h, r, G, peers, 2, 1))
                _check(L.scb_shard_finish_sort(h))
                lap("sort")
                _check(L.scb_shard_send_wait(h))
                lap("exchange_rows")
                comm.barrier()   # every rank's row writes have landed
            else:
                _check(L.scb_shard_send(h, r, G, peers, 3, 0))
                lap("exchange")
                comm.barrier()   # every rank's writes have landed
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()

            def xchg(name, ptr, row, slack=0):
                send = _dev_bytes(ptr, x.n * row, dev)
                keep[name] = comm.all_to_all_bytes(send, [c * row for c in cr], [c * row for c in rr], slack=slack)
            xchg("aux", x.aux, 8)
            xchg("packed", x.packed, prow, slack=64)
            if cfg.use_quals:
                xchg("qual1", x.qual1, L1)
            if cfg.use_names:
                send = _dev_bytes(x.names, x.name_bytes, dev)
                keep["names"] = comm.all_to_all_bytes(send, cn, rn, slack=16)
            if cfg.paired:
                xchg("seq2", x.seq2, L2)
                if cfg.use_quals:
                    xchg("qual2", x.qual2, L2)
            ev1.record()
            torch.cuda.synchronize(dev)
            ms["exchange"] = ev0.elapsed_time(ev1)
            y.aux = keep["aux"].data_ptr()
            y.packed = keep["packed"].data_ptr()
            y.qual1 = keep["qual1"].data_ptr() if "qual1" in keep else None
            y.names = keep["names"].data_ptr() if "names" in keep else None
            y.seq2 = keep["seq2"].data_ptr() if "seq2" in keep else None
            y.qual2 = keep["qual2"].data_ptr() if "qual2" in keep else None

        # ---- sort + emit of the owned slice --------------------------------------------------------------------
        if not (self.p2p and self.overlap):
            _check(L.scb_shard_import(h, C.byref(y), n_chunks))
            lap("import")
        res = ScbResult()
        _check(L.scb_shard_finish(h, C.byref(res)))
        lap("emit")
        self._keep = keep   # the received arrays back the result until the next flush
        self.stats = dict(ms=ms, rounds=rounds, n_local=n_local.value, n_recv=n_recv, n_global=n_global, n_chunks=n_chunks,
                          split=split, sent_bytes=int(sum((8 + prow + (L1 if cfg.use_quals else 0) + ((L2 + (L2 if cfg.use_quals else 0)) if cfg.paired else 0)) * c
                                                          for g, c in enumerate(cr) if g != r) + sum(c for g, c in enumerate(cn) if g != r)))
        out = FlushResult(self.t, res)
        out.n_local = n_local.value
        return out


class NcclCComm:
    """scb_comm over NCCL made by libscalce_b200_nccl.so (csrc/comm_nccl.cpp). The unique id travels through `dist`
    (any initialised torch.distributed group; a file or MPI would do as well)."""

    def __init__(self, dist, device_index):
        import os
        import torch
        from .binding import ScbComm
        here = os.path.dirname(os.path.abspath(__file__))
        self.lib = C.CDLL(os.path.join(here, "libscalce_b200_nccl.so"))
        self.lib.scb_nccl_unique_id.argtypes = [C.c_void_p]
        self.lib.scb_nccl_comm_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.POINTER(ScbComm))]
        self.lib.scb_nccl_comm_destroy.argtypes = [C.POINTER(ScbComm)]
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        idb = (C.c_uint8 * 128)()
        if self.rank == 0 and self.lib.scb_nccl_unique_id(idb) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        box = [bytes(idb)]
        dist.broadcast_object_list(box, src=0)
        idb = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self.ptr = C.POINTER(ScbComm)()
        if self.lib.scb_nccl_comm_create(idb, self.rank, self.world, device_index, C.byref(self.ptr)) != 0:
            raise RuntimeError("ncclCommInitRank failed")
        self.same_process = False

    def close(self):
        if self.ptr:
            self.lib.scb_nccl_comm_destroy(self.ptr)
            self.ptr = None


class CShardedTransform:
    """The same sharded flush through the C++ orchestrator (scb_shard_flush): the sequence of ShardedTransform.flush lives in
    the library; this class only lends it the two collectives it asks for (scb_comm: all-gather and barrier), served by any of the
    comm back-ends above. With libscalce_b200_nccl.so (an ncclComm_t wrapped into an scb_comm) no Python is involved at all."""

    PHASES = ("scan", "chunks", "resolve", "resolve_rounds", "finalize", "hist", "pack", "exchange", "import", "sort", "exchange_rows", "emit")

    def __init__(self, transform, comm, use_torch_stream=True):
        import torch
        from .binding import ALLGATHER_FN, BARRIER_FN, ScbComm, _check, load_library
        self.t, self.comm = transform, comm
        self.stats = {}
        if isinstance(comm, NcclCComm):        # a ready-made scb_comm (libscalce_b200_nccl.so): nothing of this class is on the path
            self._c, self._err = comm.ptr.contents, None
            if use_torch_stream:               # the caller's events (bench.py) then bracket the library's work and the NCCL collectives
                with torch.cuda.device(transform.cfg.device):
                    s_ = torch.cuda.current_stream().cuda_stream
                _check(load_library().scb_set_stream(transform._h, C.c_void_p(s_), 1))
            return
        dev = torch.device("cuda", transform.cfg.device)
        if use_torch_stream:
            with torch.cuda.device(transform.cfg.device):
                s = torch.cuda.current_stream().cuda_stream
            _check(load_library().scb_set_stream(transform._h, C.c_void_p(s), 1))
        self._err = None

        def allgather(ctx, send, recv, nbytes, device, stream):
            try:
                nbytes = int(nbytes)
                if device:
                    src = _dev_bytes(send, nbytes, dev)
                    if isinstance(comm, LoopbackComm):
                        torch.cuda.synchronize(dev)
                    got = comm.allgather(src)                       # [G, nbytes] on the device, ordered on the current stream
                    _dev_bytes(recv, nbytes * comm.world, dev).copy_(got.reshape(-1))
                    if isinstance(comm, LoopbackComm):
                        torch.cuda.synchronize(dev)
                else:
                    parts = comm.allgather_bytes(C.string_at(send, nbytes))
                    C.memmove(recv, b"".join(parts), nbytes * comm.world)
                return 0
            except BaseException as ex:   # noqa: BLE001 - must not propagate through the C frames
                self._err = ex
                return 1

        def barrier(ctx):
            try:
                comm.barrier()
                return 0
            except BaseException as ex:   # noqa: BLE001
                self._err = ex
                return 1
        self._cb = (ALLGATHER_FN(allgather), BARRIER_FN(barrier))      # keep the thunks alive
        self._c = ScbComm(comm.rank, comm.world, 1 if comm.same_process else 0, 0, None, self._cb[0], self._cb[1])

    def flush(self):
        from .binding import FlushResult, ScbResult, _check, load_library
        L = load_library()
        res = ScbResult()
        rc = L.scb_shard_flush(self.t._h, C.byref(self._c), C.byref(res))
        if rc != 0 and self._err is not None:
            raise self._err
        _check(rc)
        ms = (C.c_float * 12)()
        rounds = C.c_int32()
        L.scb_shard_flush_stats(self.t._h, ms, 12, C.byref(rounds))
        wall = (C.c_float * 12)()
        L.scb_shard_flush_wall(self.t._h, wall, 12)
        self.stats = dict(ms=dict(zip(self.PHASES, [float(x) for x in ms])), rounds=rounds.value, wall_ms=dict(zip(self.PHASES, [float(x) for x in wall])),
                          split=("flush chunks" if L.scb_shard_split_mode(self.t._h) == 1 else "bucket ranges"))
        out = FlushResult(self.t, res)
        out.n_local = self.t.n_local_last()
        return out

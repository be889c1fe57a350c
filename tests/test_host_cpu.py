"""CPU suite, part 2: the C-ABI library loads and exports what include/scalce_b200.h declares, the
host-side core-table compiler agrees with the oracle's automaton, and the launch-side helpers work
under a 2-process gloo group. No compute call needs a GPU here."""
import itertools
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from scalce_b200 import build
    from scalce_b200.binding import load_library
    build.build_lib()
    return load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "scalce_b200.h")).read()
    declared = set(re.findall(r"\b(scb_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from scalce_b200.binding import EXPORTS
    assert set(EXPORTS) <= declared


def test_abi_version_and_error_text(lib):
    assert lib.scb_abi_version() == 2
    assert isinstance(lib.scb_last_error(), bytes)


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from scalce_b200.binding import BoostTransform, ScbError
    with pytest.raises(ScbError) as e:
        BoostTransform(["ACGTACGT"], 50)
    assert e.value.code == -2   # SCB_ENODEVICE: fails loudly, never computes on the CPU


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "scalce_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in src.replace("oracle/ outside", ""), f"{f} references oracle/"


@pytest.mark.parametrize("cores", [
    ["ACGTACGT", "ACGTACGA", "ACGT", "TTTTTTTTTT", "ACGTACGT"],                 # duplicate: last index wins
    ["CCGTAGGT", "GGATTACA", "TTTTGGGA", "CGCGCGAT"],                          # no core starts with A
    ["".join(p) for p in itertools.product("ACGT", repeat=3)],                # dense
    ["acgtNacg", "ACGTAACG"],                                                  # case / N fold to the same core
])
def test_core_table_matches_oracle_automaton(lib, oracle_lib, cores):
    from scalce_b200.binding import table_dryrun
    t = table_dryrun(cores)
    o = oracle_lib.Oracle(cores, 40)
    assert t["n_states"] == o.n_nodes + 1
    for i in range(len(cores)):
        assert t["core_node_id"][i] == o.core_node_id(i)
    assert t["n_buckets"] == len({int(x) for x in t["core_node_id"]})


def test_core_table_random_sets(lib, oracle_lib):
    from oracle.gen_cores import make_cores
    from scalce_b200.binding import table_dryrun
    for seed in range(5):
        cores = make_cores(seed, [(5, 40), (8, 200), (13, 100), (21, 20)])
        t = table_dryrun(cores)
        o = oracle_lib.Oracle(cores, 40)
        assert t["n_states"] == o.n_nodes + 1
        assert all(t["core_node_id"][i] == o.core_node_id(i) for i in range(len(cores)))
        assert t["root_order_pos"] == t["n_buckets"]


def test_root_order_quirk_position(lib):
    from scalce_b200.binding import table_dryrun
    # nothing starts with A: aho_output meets the root first (reads.cpp:473-476)
    assert table_dryrun(["CCGT", "GGAT", "TTTT"])["root_order_pos"] == 0
    # 'A' is itself a core, nothing starts with C: the root comes right after it
    assert table_dryrun(["A", "GGAT", "TTTT"])["root_order_pos"] == 1


def test_core_file_loaders_agree(tmp_path, lib):
    """text (-P) and patterns.bin forms of the same set give the same cores in the same order."""
    from oracle.gen_cores import make_cores, write_text, write_binary, binary_order
    cores = make_cores(5, [(8, 30), (11, 20), (16, 10)])
    write_text(str(tmp_path / "c.txt"), cores)
    write_binary(str(tmp_path / "c.bin"), cores)
    # the loaders sit behind scb_create_from_file, which needs a device; check the formats themselves
    txt = open(tmp_path / "c.txt").read().split()
    assert txt == cores
    raw = open(tmp_path / "c.bin", "rb").read()
    import struct
    pos, got = 0, []
    while pos < len(raw):
        ln, cnt = struct.unpack_from("<hi", raw, pos); pos += 6
        sz = (ln + 3) // 4
        for _ in range(cnt):
            x = int.from_bytes(raw[pos:pos + sz], "little"); pos += sz
            got.append("".join("ACGT"[(x >> (2 * j)) & 3] for j in range(ln - 1, -1, -1)))
    assert got == binary_order(cores)


def test_synth_is_seeded():
    from scalce_b200 import synth
    a = synth.make_batch(100, 50, seed=3, paired=True, L2=30)
    b = synth.make_batch(100, 50, seed=3, paired=True, L2=30)
    assert (a.seq == b.seq).all() and (a.qual2 == b.qual2).all() and (a.name_off == b.name_off).all()
    assert a.qual.min() >= 33 and a.qual.max() < 33 + 80


def test_bench_algorithmic_bytes():
    sys.path.insert(0, ROOT)
    import bench
    # 3L + 2(n+1) + ceil((L-l)/4) + 1
    assert bench.algorithmic_bytes_per_read(150, 12, 12) == 3 * 150 + 2 * 13 + 35 + 1
    assert bench.algorithmic_bytes_per_read(36, 12, 12) == 141
    assert len(bench.headline_cores()) == 2048


def test_shard_bounds():
    from scalce_b200.shard import shard_bounds
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(8, 8) == list(range(9))
    b = shard_bounds(50_000_001, 8)
    assert b[0] == 0 and b[-1] == 50_000_001 and max(np.diff(b)) - min(np.diff(b)) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from scalce_b200.shard import shard_bounds, max_over_ranks, sum_over_ranks
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
r = dist.get_rank()
b = shard_bounds(1001, 2)
mine = b[r + 1] - b[r]
ms = max_over_ranks([10.0 + 5 * r, 1.0], dist)
tot = sum_over_ranks([float(mine)], dist)
assert ms == [15.0, 1.0], ms
assert tot == [1001.0], tot
dist.barrier()
dist.destroy_process_group()
print("ok", r)
'''


def test_two_rank_gloo_aggregation(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


# ---- host logic of the sharded run (scalce_b200/shard.py) -------------------------------------------
def test_balanced_split_properties():
    from scalce_b200.shard import balanced_split
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        for _ in range(20):
            ncols = int(rng.integers(1, 400))
            hist = rng.integers(0, 1000, size=ncols)
            if rng.random() < 0.3:
                hist[int(rng.integers(0, ncols))] += 10 ** 6        # one dominant bucket (e.g. the root)
            s = balanced_split(hist, world)
            assert len(s) == world + 1 and s[0] == 0 and s[-1] == ncols
            assert all(a <= b for a, b in zip(s[:-1], s[1:]))
    # even histogram splits evenly; a bucket is never divided
    assert balanced_split([10] * 8, 4) == [0, 2, 4, 6, 8]
    assert balanced_split([0, 0, 100, 0], 2) in ([0, 2, 4], [0, 3, 4])


def _ref_chunk_ids(sizes, B):
    """compress.cpp:702, 708-713 over the whole input: chunk id of every read."""
    out, tot, c = [], 0, 0
    for s in sizes:
        out.append(c)
        tot += s
        if tot >= B:
            c += 1
            tot = 0
    return out, c + (1 if tot > 0 else 0)


def _shard_sizes_host(sizes, B, carry, chunk):
    """Host restatement of scb_shard_sizes (chunk_bounds_carry_k): returns ids, carry_out, chunk_out."""
    ids = []
    for s in sizes:
        ids.append(chunk)
        carry += s
        if carry >= B:
            chunk += 1
            carry = 0
    return ids, carry, chunk


def test_chunk_chain_equals_global_numbering():
    from scalce_b200.shard import chain_chunks, shard_bounds

    class FakeComm:
        def __init__(self, world):
            self.world, self.rank, self.box = world, 0, None
    rng = np.random.default_rng(5)
    for world in (1, 2, 5):
        sizes = rng.integers(100, 400, size=1000).tolist()
        B = 7000
        want_ids, want_n = _ref_chunk_ids(sizes, B)
        bd = shard_bounds(len(sizes), world)
        # ranks take turns in rank order: emulate the chain sequentially
        carry, chunk, got = 0, 0, []
        for g in range(world):
            ids, carry, chunk = _shard_sizes_host(sizes[bd[g]:bd[g + 1]], B, carry, chunk)
            got += ids
        assert got == want_ids
        assert chunk + (1 if carry > 0 else 0) == want_n


_GLOO_SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from scalce_b200.shard import TorchComm, chain_chunks, balanced_split, shard_bounds
rank = int(sys.argv[1]); world = 2
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
comm = TorchComm(dist, torch.device("cpu"))
# host metadata collectives
assert comm.allgather_host([rank, 10 + rank]) == [[0, 10], [1, 11]]
assert comm.bcast_host([7 * (rank + 1), 3], src=1) == [14, 3]
# chunk chain: rank-order hand-off of (carry, chunk)
rng = np.random.default_rng(11)
sizes = rng.integers(100, 400, size=600).tolist(); B = 5000
bd = shard_bounds(len(sizes), world)
lay = {{}}
def sizes_fn(carry, chunk):
    mine = sizes[bd[rank]:bd[rank + 1]]
    starts = []                                   # local indices at which a new chunk starts (may equal len(mine))
    c0 = chunk
    for i, s in enumerate(mine):
        carry += s
        if carry >= B: chunk += 1; carry = 0; starts.append(i + 1)
    n = len(mine)
    lay["v"] = [c0, len(starts), n, starts[0] if starts else n, (n - starts[-1]) if starts else 0]   # = scb_shard_chunk_layout
    return carry, chunk
n_chunks = chain_chunks(sizes_fn, comm)
# flush-chunk ownership: every rank derives the same owners from the all-gathered layouts (scb_shard_chunk_owners: host arithmetic)
import ctypes as C
from scalce_b200.binding import load_library
L = load_library()
lays = np.array(comm.allgather_host(lay["v"]), dtype=np.int64)
owner = (C.c_int32 * n_chunks)(); ml = C.c_int64()
assert L.scb_shard_chunk_owners(lays.ctypes.data_as(C.POINTER(C.c_int64)), world, n_chunks, owner, C.byref(ml)) == 0
owners = list(owner)
assert owners == sorted(owners) and set(owners) <= {{0, 1}}
assert comm.allgather_host(owners) == [owners, owners]          # identical on both ranks
# the owner of a chunk holds part of it, and the loads add up
chunk_of, c, tot_ = [], 0, 0
for s_ in sizes:
    chunk_of.append(c); tot_ += s_
    if tot_ >= B: c += 1; tot_ = 0
load = [0, 0]
for i, c_ in enumerate(chunk_of):
    load[owners[c_]] += 1
for c_ in range(n_chunks):
    members = [i for i, x in enumerate(chunk_of) if x == c_]
    assert any(bd[owners[c_]] <= i < bd[owners[c_] + 1] for i in members), (c_, owners[c_])
assert sum(load) == len(sizes) and max(load) == ml.value, (load, ml.value)
tot, c = 0, 0
for s in sizes:
    tot += s
    if tot >= B: c += 1; tot = 0
assert n_chunks == c + (1 if tot > 0 else 0), (n_chunks, c, tot)
# histogram all-gather -> identical split on every rank
hist = torch.tensor([5, 0, 9, 1] if rank == 0 else [1, 2, 3, 40], dtype=torch.int32)
g = comm.allgather(hist).to(torch.int64).sum(0).numpy()
split = balanced_split(g, world)
assert split[0] == 0 and split[-1] == 4
# payload all-to-all (bytes, destination-major) with slack
send = torch.arange(10, dtype=torch.uint8) + 100 * rank
cnt = [3, 7] if rank == 0 else [6, 4]
mat = comm.allgather_host(cnt)
recv_cnt = [mat[s][rank] for s in range(world)]
buf = comm.all_to_all_bytes(send, cnt, recv_cnt, slack=8)
got = buf[:sum(recv_cnt)].tolist()
want = (list(range(0, 3)) + list(range(100, 106))) if rank == 0 else (list(range(3, 10)) + list(range(106, 110)))
assert got == want, (got, want)
assert buf.numel() == sum(recv_cnt) + 8
comm.barrier()
dist.destroy_process_group()
print("ok", rank, split)
'''


def test_two_rank_gloo_shard_plumbing(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "ws.py"
    script.write_text(_GLOO_SHARD_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
    assert outs[0].split("ok 0")[1].strip() == outs[1].split("ok 1")[1].strip()   # same split everywhere


def test_aux_word_layout_matches_header():
    # include/scalce_b200.h documents the exchange word: rank | end << 24 (12 bits) | name length << 36 | chunk << 44
    hdr = open(os.path.join(ROOT, "include", "scalce_b200.h")).read()
    cuh = open(os.path.join(ROOT, "scalce_b200", "csrc", "shard.cuh")).read()
    assert "end << 24" in hdr and "<< 36" in hdr and "<< 44" in hdr
    assert "<< 24" in cuh and "<< 36" in cuh and "<< 44" in cuh


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in scalce_b200/binding.py against the C compiler's view of include/scalce_b200.h."""
    import ctypes as C
    from scalce_b200 import binding as B
    structs = {"scb_config": B.ScbConfig, "scb_batch": B.ScbBatch, "scb_result": B.ScbResult,
               "scb_shard_xfer": B.ScbShardXfer, "scb_shard_peer": B.ScbShardPeer}
    fields = {"scb_config": ["read_length", "use_names", "paired", "use_quals", "device", "bucket_set_bytes", "emit_merged"],
              "scb_batch": ["n", "seq1", "qual1", "names", "name_off", "seq2", "qual2", "location"],
              "scb_result": ["n_reads", "n_chunks", "n_buckets_nonempty", "data", "chunk_off", "merged", "merged_size", "bucket_id", "core_idx",
                             "end", "chunk", "perm", "device_ms"],
              "scb_shard_xfer": ["n", "name_bytes", "aux", "packed", "qual1", "names", "seq2", "qual2", "cnt_reads", "cnt_name_bytes", "packed_row_bytes"],
              "scb_shard_peer": ["aux", "packed", "qual1", "names", "seq2", "qual2", "row_off", "name_off"]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "scalce_b200.h")}"', 'int main(void) {']
    for s, fl in fields.items():
        src.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for f in fl:
            src.append(f'  printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    src += ['  printf("SCB_N_STREAMS %d\\n", (int)SCB_N_STREAMS);', '  printf("SCB_ABI_VERSION %d\\n", (int)SCB_ABI_VERSION);', '  return 0;', '}']
    c = tmp_path / "abi.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(c), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines())
    for s, cls in structs.items():
        assert int(out[s]) == C.sizeof(cls), f"sizeof({s}): C {out[s]} vs ctypes {C.sizeof(cls)}"
        for f in fields[s]:
            assert int(out[f"{s}.{f}"]) == getattr(cls, f).offset, f"offsetof({s}, {f})"
    assert int(out["SCB_N_STREAMS"]) == B.N_STREAMS
    from scalce_b200.binding import load_library
    assert int(out["SCB_ABI_VERSION"]) == load_library().scb_abi_version()


# ---- C++ host side (scalce_b200/host/scb_boost.cpp): parse + payload preparation, no GPU involved -------------------
def _run_host_tool_dump(tmp_path, b, paired=False, extra=()):
    import subprocess
    from scalce_b200 import build as bld, synth
    tool = bld.build_host_tool()
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(b, f1, f2 if paired else None)
    d = tmp_path / "soa"
    d.mkdir()
    cmd = [tool, f1] + (["-r", f2] if paired else []) + ["--dump-soa", str(d)] + list(extra)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    meta = dict(line.split() for line in (d / "meta.txt").read_text().splitlines())
    arr = {k: np.fromfile(d / f"{k}.bin", dtype=np.uint8) for k in ("seq1", "qual1", "names", "seq2", "qual2")}
    arr["name_off"] = np.fromfile(d / "name_off.bin", dtype=np.int64)
    return {k: int(v) for k, v in meta.items()}, arr


def test_host_tool_parse_and_payload_single_end(tmp_path):
    """FASTQ -> SoA exactly as the Python host side builds it: names without '@', q - offset with 0 under 'N'
    (qualities.cpp:177-204 as restated by oracle.quality_payload), several batches concatenated."""
    from oracle import oracle as orc
    from scalce_b200 import synth
    b = synth.make_batch(5000, 100, seed=11, n_frac=0.01, lower_frac=0.05)
    meta, a = _run_host_tool_dump(tmp_path, b, extra=("--batch", "777"))
    assert meta == {"n": 5000, "L1": 100, "L2": 0, "offset1": 33, "offset2": 0}
    assert np.array_equal(a["seq1"].reshape(5000, 100), b.seq)
    assert np.array_equal(a["qual1"].reshape(5000, 100), orc.quality_payload(b.qual, b.seq, orc.detect_phred_offset(b.qual)))
    assert np.array_equal(a["names"], b.names) and np.array_equal(a["name_off"], b.name_off)


def test_host_tool_paired_phred64_and_name_truncation(tmp_path):
    from oracle import oracle as orc
    from scalce_b200 import synth
    b = synth.make_batch(1200, 75, seed=12, paired=True, L2=50)
    # phred+64 qualities on mate 2 only: the offsets are detected per file (compress.cpp:561-574)
    b.qual2 = (b.qual2 - 33 + 64).astype(np.uint8)
    # names with a description after a space: only the part before it is kept (names.cpp:48-62)
    full = []
    for i in range(b.n):
        nm = b.names[b.name_off[i]:b.name_off[i + 1]].tobytes()
        full.append(nm + (b" extra field %d" % i if i % 3 == 0 else b""))
    kept_names, kept_off = b.names.copy(), b.name_off.copy()
    b.names = np.frombuffer(b"".join(full), dtype=np.uint8).copy()
    off = np.zeros(b.n + 1, dtype=np.int64)
    np.cumsum([len(x) for x in full], out=off[1:])
    b.name_off = off
    meta, a = _run_host_tool_dump(tmp_path, b, paired=True)
    assert meta == {"n": 1200, "L1": 75, "L2": 50, "offset1": 33, "offset2": 64}
    assert np.array_equal(a["names"], kept_names) and np.array_equal(a["name_off"], kept_off)
    assert np.array_equal(a["seq2"].reshape(1200, 50), b.seq2)
    assert np.array_equal(a["qual1"].reshape(1200, 75), orc.quality_payload(b.qual, b.seq, 33))
    assert np.array_equal(a["qual2"].reshape(1200, 50), orc.quality_payload(b.qual2, b.seq2, 64))


def test_host_tool_rejects_ragged_reads(tmp_path):
    import subprocess
    from scalce_b200 import build as bld
    tool = bld.build_host_tool()
    f = tmp_path / "bad.fastq"
    f.write_bytes(b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nACGT\n+\nIIII\n@c\nACGTACGT\n+\nIIIIIIII\n")
    (tmp_path / "o").mkdir()
    r = subprocess.run([tool, str(f), "--dump-soa", str(tmp_path / "o")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"(ERROR)" in r.stderr          # the reference's convention: message + exit(1)


def test_emit_record_host_device_code_against_per_base_reference(tmp_path):
    """scalce_b200/csrc/emit_record.h is compiled into the (opt-in) stream-1 kernel AND is plain C++: check the record it
    writes (rotation around the core, zero padding, 1- or 2-byte end marker, any byte alignment) base by base, under
    the address and undefined-behaviour sanitizers, for read lengths around every word boundary."""
    import subprocess
    src = tmp_path / "er_test.cpp"
    hdr = os.path.join(ROOT, "scalce_b200", "csrc", "emit_record.h")
    src.write_text(r'''
#include "%s"
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
using namespace scb;
int main() {
    std::mt19937 rng(7);
    long cases = 0;
    for (int L1 : {16, 17, 31, 32, 33, 36, 47, 48, 64, 100, 150, 151, 250, 255, 256, 300, 2047}) {
        const int PW = (L1 + 15) / 16, sz_meta = L1 > 255 ? 2 : 1;
        for (int rep = 0; rep < 300; rep++) {
            std::vector<int> base(L1);
            for (auto &b : base) b = rng() & 3;
            std::vector<uint32_t> row(PW + 2, 0);                    // the row + the two pad words the code may read
            for (int q = 0; q < L1; q++) row[q >> 4] |= (uint32_t)base[q] << (30 - 2 * (q & 15));
            int lv = rep %% 7 == 0 ? 0 : (int)(rng() %% (L1 < 32 ? L1 : 32)) + 1;
            int end = lv == 0 ? 0 : lv + (int)(rng() %% (L1 - lv + 1));
            if (rep %% 11 == 0 && lv) end = L1;
            if (rep %% 13 == 0 && lv) end = lv;
            std::vector<int> outb;
            for (int q = end; q < L1; q++) outb.push_back(base[q]);          // output_read, reads.cpp:432-461
            for (int q = 0; q < end - lv; q++) outb.push_back(base[q]);
            const int total = L1 - lv, nbytes = (total + 3) / 4;
            std::vector<uint8_t> want(nbytes + sz_meta, 0), got(nbytes + sz_meta + 8, 0xAA);
            for (int j = 0; j < total; j++) want[j >> 2] |= outb[j] << (6 - 2 * (j & 3));
            want[nbytes] = end & 0xff;
            if (sz_meta > 1) want[nbytes + 1] = (end >> 8) & 0xff;
            const int sz = emit_record(row.data(), L1, lv, end, sz_meta, got.data() + 3);
            if (sz != nbytes + sz_meta || memcmp(want.data(), got.data() + 3, sz) != 0 || got[3 + sz] != 0xAA || got[2] != 0xAA) {
                printf("MISMATCH L1 %%d lv %%d end %%d\n", L1, lv, end);
                return 1;
            }
            cases++;
        }
    }
    printf("ok %%ld\n", cases);
    return 0;
}
''' % hdr)
    exe = tmp_path / "er_test"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fsanitize=address,undefined", "-o", str(exe), str(src)], check=True)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and r.stdout.startswith(b"ok"), (r.stdout + r.stderr).decode()


def _bench_dry(*args):
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    return json.loads(r.stdout.decode().strip().splitlines()[-1])


def test_bench_native_arm_control_flow_and_json_contract():
    """bench.py's native arm with the device and the library faked (tests/bench_dry_run.py): every key of the bench
    contract is present and the end-to-end arm, its pipelined variant and the per-stage figures are assembled."""
    d = _bench_dry("--reads", "20000", "--steps", "2", "--warmup", "1", "--no-cpu", "--e2e-depth", "2")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["config"]["workload"] and d["dtype"] == "u8" and d["vs_baseline"] is None and d["gpu_launches"] > 0
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "per_stage", "stage_ms"):
        assert k in rf, k
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 20000 * 100 * 0 + 20000 * 150 * 2 + 20000 * 13 + 20001 * 8 and e["d2h_bytes_per_step"] == 20 + 60 + 200 + 16
    assert e["pipelined"]["depth"] == 2 and e["pipelined"]["slots_agree"] and "serial" in e
    assert e["steps"] >= 2 and "ms_per_step_min" in e["serial"]
    assert d["config"]["name"] == "c2" and d["config"]["core_set"]["cores"] == 2048 and d["scaling"] == "weak"


def test_bench_named_configs():
    """--config c3 / c4 / c5 (BASELINE configs[2..4]): paired input with both mates in the byte counts, the 36 bp shape with the
    total split over the GPUs (strong scaling), 250 bp with high-entropy qualities."""
    import bench
    d = _bench_dry("--config", "c3", "--reads", "3000", "--steps", "1", "--warmup", "1", "--no-cpu", "--e2e-depth", "1", "--e2e-steps", "1")
    assert d["config"]["paired"] and d["config"]["mate2_length"] == 150 and "paired-end" in d["config"]["workload"]
    assert d["e2e"]["h2d_bytes_per_step"] == 3000 * 150 * 4 + 3000 * 13 + 3001 * 8
    assert d["e2e"]["d2h_bytes_per_step"] == 20 + 60 + 200 + 16 + 80 + 200
    assert d["pipeline_roofline"]["bytes_per_read"] == bench.algorithmic_bytes_per_read(150, 13, 8, True, 150)
    d = _bench_dry("--config", "c4", "--reads", "5000", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e")
    assert d["config"]["read_length"] == 36 and d["scaling"] == "strong" and d["config"]["flushes_per_step"] == 1
    d = _bench_dry("--config", "c5", "--reads", "2000", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-e2e")
    assert d["config"]["read_length"] == 250 and d["config"]["high_entropy_qualities"]
    assert bench.CONFIGS["c4"]["total"] == 500_000_000 and bench.CONFIGS["c3"]["reads"] * 8 == 200_000_000 and bench.CONFIGS["c5"]["reads"] * 8 == 100_000_000


def test_bench_native_arm_cpu_baseline_fields():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")):
        pytest.skip("oracle/_ref not built")
    d = _bench_dry("--reads", "20000", "--steps", "1", "--warmup", "1", "--no-e2e", "--cpu-sample", "20000")
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] >= c["value_1thread"] * 0.5 and c["unit"] == "reads/s" and c["sample"]
    assert d["e2e"] is None


def test_stream1_checker_against_the_oracle():
    """tests/util.py::stream1_window_matches (the torch check the full-size GPU test applies to stream 1 at 50M reads) must accept
    the oracle's stream 1 and reject a corrupted one. Output order recovered from the (unique) names in the oracle's stream 0."""
    import torch
    from tests import util
    cores, b, q1, q2, _ = util.make_case(6000, 100, seed=33)
    o = util.run_oracle(cores, b, q1, q2)
    assert o.n_chunks == 1
    s0 = np.frombuffer(o.stream(0, 0), dtype=np.uint8)
    name_to_idx = {b.names[b.name_off[i]:b.name_off[i + 1]].tobytes(): i for i in range(b.n)}
    perm, p = [], 0
    while p < s0.size:
        nl = int(s0[p])
        perm.append(name_to_idx[s0[p + 1:p + 1 + nl].tobytes()])
        p += 1 + nl
    perm = torch.tensor(perm, dtype=torch.int64)
    assert sorted(perm.tolist()) == list(range(b.n))
    d = o.debug()
    lens = np.array([len(c) for c in cores] + [0])
    lv = torch.from_numpy(lens[np.where(d["core"] >= 0, d["core"], len(cores))].astype(np.int64))
    end = torch.from_numpy(d["end"].astype(np.int64))
    s1 = torch.from_numpy(np.frombuffer(o.stream(1, 0), dtype=np.uint8).copy())
    seq = torch.from_numpy(b.seq)
    assert util.stream1_window_matches(s1, seq, perm, lv, end, 100, 0, b.n)
    assert util.stream1_window_matches(s1, seq, perm, lv, end, 100, 2500, 1000)
    bad = s1.clone()
    bad[len(bad) // 2] ^= 0x10
    assert not util.stream1_window_matches(bad, seq, perm, lv, end, 100, 0, b.n)


def test_oracle_streaming_assign_equals_full_oracle():
    """orc_assign (the payload-free form used to check all 50 M reads of the bench workload) gives the full oracle's per-read
    arrays, fed in pieces, multi-chunk, with and without names."""
    from tests import util
    from oracle import oracle as orc
    for use_names in (True, False):
        cores, b, q1, q2, _ = util.make_case(30000, 100, seed=401)
        o = util.run_oracle(cores, b, q1, q2, use_names=use_names, bucket_set_bytes=1 << 20)
        want = o.debug()
        assert o.n_chunks > 3
        s = orc.Oracle(cores, 100, use_names=use_names, bucket_set_bytes=1 << 20)
        got = {k: [] for k in want}
        for a, z in ((0, 7), (7, 12001), (12001, 30000)):
            d = s.assign(b.seq[a:z], b.name_off[a:z + 1] if use_names else None)
            for k in got:
                got[k].append(d[k])
        for k in want:
            assert np.array_equal(want[k], np.concatenate(got[k])), k
        assert s.unbucketed == o.unbucketed


def test_vectorised_core_set_generator():
    from scalce_b200 import synth
    from scalce_b200.binding import table_dryrun
    spec = [(8, 300), (10, 5000), (13, 20000)]
    c = synth.make_core_set(spec, seed=3)
    assert len(c) == 25300 and len(set(c)) == 25300
    assert [len(x) for x in c[:300]] == [8] * 300 and len(c[-1]) == 13 and set("".join(c[:50])) <= set("ACGT")
    assert c == synth.make_core_set(spec, seed=3) and c != synth.make_core_set(spec, seed=4)
    d = table_dryrun(c)
    assert d["n_buckets"] == 25300
    assert synth.make_core_set([(3, 1000)], seed=1).__len__() == 64      # capped at 4^length


def test_chunk_owner_rule_on_simulated_shards():
    """scb_shard_chunk_owners (host arithmetic of the sharded flush): whole flush chunks to ranks. Random shard and chunk
    boundaries; the owners must not decrease, must touch their chunk, and the loads must add up to the input."""
    import ctypes as C
    from scalce_b200.binding import load_library
    L = load_library()
    rng = np.random.default_rng(11)
    for trial in range(300):
        G = int(rng.integers(1, 9))
        n = int(rng.integers(1, 5000))
        n_chunks = int(rng.integers(1, 40))
        sb = np.sort(rng.integers(0, n + 1, size=G - 1)) if G > 1 else np.zeros(0, dtype=np.int64)
        shard = np.concatenate([[0], sb, [n]]).astype(np.int64)                      # shard g = [shard[g], shard[g+1])
        cb = np.sort(rng.choice(np.arange(1, n), size=min(n_chunks - 1, n - 1), replace=False)) if n > 1 else np.zeros(0, dtype=np.int64)
        chunk_start = np.concatenate([[0], cb]).astype(np.int64)                      # chunk c = [chunk_start[c], chunk_start[c+1])
        n_chunks = len(chunk_start)
        chunk_of = np.searchsorted(chunk_start, np.arange(n), side="right") - 1
        lay = np.zeros((G, 5), dtype=np.int64)
        cur_chunk = 0
        for g in range(G):
            a, z = int(shard[g]), int(shard[g + 1])
            # chunk id the shard starts in (a chunk starting exactly at the shard's first read was announced by the rank before:
            # its flush fell on its last read)
            cin = int(chunk_of[a]) if a < n else n_chunks - 1
            if a == z:
                lay[g] = [cin, 0, 0, 0, 0]
                continue
            starts = [int(x) - a for x in chunk_start if a < x < z]                # chunks that start inside the shard, local indices
            # a chunk starting exactly at z (the shard's end) is announced by this rank as a boundary at local index n
            if z < n and z in chunk_start:
                starts.append(z - a)
            head = starts[0] if starts else z - a
            tail = (z - a) - starts[-1] if starts else 0
            lay[g] = [cin, len(starts), z - a, head, tail]
        owner = (C.c_int32 * n_chunks)()
        ml = C.c_int64()
        assert L.scb_shard_chunk_owners(lay.ctypes.data_as(C.POINTER(C.c_int64)), G, n_chunks, owner, C.byref(ml)) == 0
        ow = np.array(owner[:])
        assert (np.diff(ow) >= 0).all(), (trial, ow)
        load = np.zeros(G, dtype=np.int64)
        rank_of = np.searchsorted(shard, np.arange(n), side="right") - 1
        rank_of = np.minimum(rank_of, G - 1)
        for c in range(n_chunks):
            members = np.nonzero(chunk_of == c)[0]
            assert ow[c] in set(rank_of[members].tolist()), (trial, c, ow[c])    # the owner holds part of its chunk
            load[ow[c]] += len(members)
        assert load.sum() == n and load.max() == ml.value, (trial, load, ml.value)

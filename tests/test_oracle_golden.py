"""CPU suite, part 1: the oracle (oracle/scalce_oracle.c) against fixtures the UNMODIFIED reference
CLI produced (tests/golden/*.npz, generator tests/make_golden.py), byte for byte."""
import ast
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.gen_cores import make_cores
from scalce_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    meta = ast.literal_eval(z["meta"].tobytes().decode())
    return z, meta


def _inputs(z, meta):
    if meta["store"] == "full":
        cores = z["cores"].tobytes().decode().split("\n")
        b = synth.FastqBatch(z["seq"], z["qual"], z["names"], z["name_off"],
                             z["seq2"] if meta["paired"] else None, z["qual2"] if meta["paired"] else None)
        return cores, b
    cores = make_cores(meta["seed"], [tuple(x) for x in meta["spec"]])
    b = synth.make_batch(meta["n"], meta["L"], seed=meta["seed"], paired=meta["paired"], L2=meta["L2"] or None, lower_frac=meta["lower"])
    synth.plant_cores(b, cores, seed=meta["seed"] + 1, frac=0.5)
    h = hashlib.sha256()
    h.update("\n".join(cores).encode())
    for a in (b.seq, b.qual, b.names, b.name_off, b.seq2, b.qual2):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    if h.hexdigest() != meta["input_sha"]:
        pytest.skip("numpy generator stream differs from the one the fixture was made with")
    return cores, b


def oracle_files(cores, b, meta):
    off = orc.detect_phred_offset(b.qual)
    q1 = orc.quality_payload(b.qual, b.seq, off)
    q2 = orc.quality_payload(b.qual2, b.seq2, orc.detect_phred_offset(b.qual2)) if meta["paired"] else None
    bucket = {"4G": 4 << 30, "1M": 1 << 20}[meta["bucket"]]
    o = orc.Oracle(cores, meta["L"], meta["L2"], use_names=meta["use_names"], paired=meta["paired"], bucket_set_bytes=bucket)
    o.submit(b.seq, q1, b.names, b.name_off, b.seq2, q2)
    o.finish()
    out = {}
    for mate in range(1 + int(meta["paired"])):
        fn, fr, fq = orc.assemble_container(o.stream(3), o.stream(0), o.stream(1), o.stream(2), cores, meta["L"], off,
                                            use_names=meta["use_names"], library=b"lib", paired=meta["paired"], mate=mate,
                                            reads2=o.stream(4), quals2=o.stream(5), L2=meta["L2"])
        out[f"{mate + 1}n"], out[f"{mate + 1}r"], out[f"{mate + 1}q"] = fn, fr, fq
    return o, out


def test_fixtures_present():
    assert len(GOLD) >= 7


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_cli(path, oracle_lib):
    z, meta = _load(path)
    assert meta["roundtrip_ok"], "fixture was not round-tripped by the reference decompressor"
    cores, b = _inputs(z, meta)
    o, files = oracle_files(cores, b, meta)
    for k, data in files.items():
        assert len(data) == meta["sizes"][k], f"{k}: size {len(data)} != reference {meta['sizes'][k]}"
        assert hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{k}: bytes differ from the reference CLI output"
        if meta["store"] == "full":
            assert data == z["out_" + k].tobytes()
    if meta["bucket"] == "1M":
        assert o.n_chunks > 1   # the multi-chunk merge path (compress.cpp:68-198) is what these fixtures pin


@pytest.mark.parametrize("n,L", [(3000, 80), (300, 2047), (300, 2498)])
def test_oracle_against_live_reference_cli(tmp_path, oracle_lib, n, L):
    """When oracle/_ref/scalce exists (build container), run it now on a fresh seed. 2498 bases is the longest read the
    reference's line buffer holds (fgets into MAXLINE = 2500 bytes, const.h:87): the 2-byte end marker at its largest."""
    if not os.path.exists(orc.REF_CLI):
        pytest.skip("reference CLI not built here")
    from oracle.gen_cores import write_text
    spec = [(8, 128), (9, 64), (11, 32)]
    cores = make_cores(911, spec)
    b = synth.make_batch(n, L, seed=911, lower_frac=0.02)
    synth.plant_cores(b, cores, seed=912, frac=0.5)
    d = str(tmp_path)
    write_text(d + "/cores.txt", cores)
    synth.write_fastq(b, d + "/in_1.fastq")
    orc.run_reference_cli(d + "/in_1.fastq", d + "/ref", d + "/cores.txt", tmpdir=d + "/tmp")
    meta = dict(L=L, L2=0, paired=False, use_names=True, bucket="4G")
    _, files = oracle_files(cores, b, meta)
    for ext in "nrq":
        assert files["1" + ext] == open(f"{d}/ref_1.scalce{ext}", "rb").read()


def test_reference_harness_thread_loop_conserves_payload(tmp_path):
    """oracle/_ref/libref_harness.so (bench.py's CPU arm): the -T loop (refh_run_mt) must move the same reads as the
    single-threaded loop - its order is not deterministic, but the bytes of the names and quality files and the
    number of reads per file set are."""
    import ctypes as C
    import numpy as np
    import bench
    harness = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_harness.so")
    if not os.path.exists(harness):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cores = bench.headline_cores()
    n, L = 20000, 100
    seq, qual, names, off, _, _ = bench.synth_host_sample(n, L, seed=3)
    H = C.CDLL(harness)
    H.refh_init.restype = C.c_double
    H.refh_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
    H.refh_run.restype = C.c_double
    H.refh_run.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    H.refh_run_mt.restype = C.c_double
    H.refh_run_mt.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    cf = tmp_path / "cores.txt"
    cf.write_text("\n".join(cores) + "\n")
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    sizes = []
    for mode, threads in (("st", 1), ("mt", 4)):
        d = tmp_path / mode
        d.mkdir()
        H.refh_init(str(cf).encode(), L, 0, 0, 1, 4 << 30)
        nch = C.c_int()
        if mode == "st":
            H.refh_run(n, p(seq), p(qual), p(names), p(off), None, None, 33, str(d).encode(), C.byref(nch), None, None)
        else:
            H.refh_run_mt(n, p(seq), p(qual), p(names), p(off), None, None, 33, str(d).encode(), C.byref(nch), threads)
        assert nch.value == 1
        sizes.append([os.path.getsize(d / f"t_000_{k}.tmp") for k in range(4)])
    st, mt = sizes
    assert st[0] == mt[0] == int(off[n]) + n          # names: length byte + name per read
    assert st[2] == mt[2] == n * L                    # qualities
    assert st[1] == mt[1]                             # packed reads: the chosen core's LENGTH does not depend on the tie-break


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cpp_host_stage_plus_oracle_matches_reference_cli(path, oracle_lib, tmp_path):
    """The C++ host side (scalce_b200/host/scb_boost --dump-soa: FASTQ parse, phred offset, name and quality payloads)
    feeding the oracle must give the files the UNMODIFIED reference CLI wrote for the same FASTQ: the host stages of the
    drop-in are pinned to the reference without a GPU (the transform in between is what the -m gpu tests pin)."""
    import subprocess
    from scalce_b200 import build as bld
    z, meta = _load(path)
    cores, b = _inputs(z, meta)
    tool = bld.build_host_tool()
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(b, f1, f2 if meta["paired"] else None)
    d = tmp_path / "soa"
    d.mkdir()
    cmd = [tool, f1] + (["-r", f2] if meta["paired"] else []) + ["--dump-soa", str(d), "--batch", "1000"] + ([] if meta["use_names"] else ["-n"])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    info = dict(line.split() for line in (d / "meta.txt").read_text().splitlines())
    n, L, L2 = int(info["n"]), int(info["L1"]), int(info["L2"])
    assert (n, L, L2) == (meta["n"] if "n" in meta else b.n, meta["L"], meta["L2"] or 0)
    rd = lambda k, dt=np.uint8: np.fromfile(d / f"{k}.bin", dtype=dt)
    seq, q1 = rd("seq1").reshape(n, L), rd("qual1").reshape(n, L)
    seq2 = rd("seq2").reshape(n, L2) if meta["paired"] else None
    q2 = rd("qual2").reshape(n, L2) if meta["paired"] else None
    names, name_off = (rd("names"), rd("name_off", np.int64)) if meta["use_names"] else (b.names, b.name_off)
    bucket = {"4G": 4 << 30, "1M": 1 << 20}[meta["bucket"]]
    o = orc.Oracle(cores, L, L2, use_names=meta["use_names"], paired=meta["paired"], bucket_set_bytes=bucket)
    o.submit(seq, q1, names, name_off, seq2, q2)
    o.finish()
    for mate in range(1 + int(meta["paired"])):
        fn, fr, fq = orc.assemble_container(o.stream(3), o.stream(0), o.stream(1), o.stream(2), cores, L, int(info["offset1"]),
                                            use_names=meta["use_names"], library=b"lib", paired=meta["paired"], mate=mate,
                                            reads2=o.stream(4), quals2=o.stream(5), L2=L2)
        for ext, data in (("n", fn), ("r", fr), ("q", fq)):
            k = f"{mate + 1}{ext}"
            assert hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{k}: bytes differ from the reference CLI output"


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cpp_container_assembly_matches_reference_cli(path, oracle_lib, tmp_path):
    """The other host stage of the drop-in, after the transform: scb_boost --assemble writes .scalce{n,r,q} from the merged
    streams as combine_and_compress_with_split does in raw mode (compress.cpp:262-379). Fed with the oracle's streams it must
    reproduce the files of the unmodified reference CLI. No GPU involved."""
    import subprocess
    from scalce_b200 import build as bld
    z, meta = _load(path)
    cores, b = _inputs(z, meta)
    o, _ = oracle_files(cores, b, meta)
    tool = bld.build_host_tool()
    d = tmp_path / "streams"
    d.mkdir()
    for k in range(6 if meta["paired"] else 4):
        (d / f"merged_{k}.tmp").write_bytes(o.stream(k))
    (tmp_path / "cores.txt").write_text("\n".join(cores) + "\n")
    off = orc.detect_phred_offset(b.qual)
    cmd = [tool, "--assemble", str(d), "--container", str(tmp_path / "out"), "-P", str(tmp_path / "cores.txt"), "--L1", str(meta["L"]),
           "--offset", str(off), "--library", "lib"] + (["--L2", str(meta["L2"])] if meta["paired"] else []) + ([] if meta["use_names"] else ["-n"])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    for mate in range(1 + int(meta["paired"])):
        for ext in "nrq":
            k = f"{mate + 1}{ext}"
            data = (tmp_path / f"out_{mate + 1}.scalce{ext}").read_bytes()
            assert len(data) == meta["sizes"][k] and hashlib.sha256(data).hexdigest() == meta["sha"][k], f"{k}: differs from the reference CLI output"


def test_oracle_inverse_matches_reference_decompressor(tmp_path, oracle_lib):
    """orc_inverse (oracle restatement of decompress.cpp:331-352, the checker of the device-side inverse scb_inverse_reads)
    against the UNMODIFIED reference: compress with the CLI (-c no -A), decompress with the CLI, and the oracle must rebuild
    the same read and quality lines, in the same order, from the .scalcer / .scalceq files."""
    if not os.path.exists(orc.REF_CLI):
        pytest.skip("reference CLI not built here")
    from oracle.gen_cores import write_text
    for seed, L, kw in ((931, 80, {}), (932, 300, {}), (933, 36, dict(lower_frac=0.05))):
        cores = make_cores(seed, [(8, 128), (9, 64), (11, 32)])
        b = synth.make_batch(2500, L, seed=seed, n_frac=0.01, **kw)
        synth.plant_cores(b, cores, seed=seed + 1, frac=0.5)
        d = str(tmp_path / f"c{seed}")
        os.makedirs(d)
        write_text(d + "/cores.txt", cores)
        synth.write_fastq(b, d + "/in_1.fastq")
        orc.run_reference_cli(d + "/in_1.fastq", d + "/ref", d + "/cores.txt", bucket="1M", tmpdir=d + "/tmp")
        orc.run_reference_decompress(d + "/ref_1.scalcen", d + "/rt", d + "/cores.txt")
        rt = open(d + "/rt_1.fastq", "rb").read().split(b"\n")
        want_seq, want_q = rt[1::4], rt[3::4]
        fr = open(d + "/ref_1.scalcer", "rb").read()
        fq = open(d + "/ref_1.scalceq", "rb").read()
        assert fr[:8] == orc.MAGIC and fq[:8] == orc.MAGIC
        phred = int(np.frombuffer(fq[8:16], dtype=np.int64)[0])
        stream, sc, sr = orc.split_reads_container(fr[16:], cores, L)
        o = orc.Oracle(cores, L)
        seq, q = o.inverse(stream, sc, sr, quals=fq[16:], phred=phred)
        assert seq.shape[0] == b.n == len(want_seq) - (1 if want_seq[-1] == b"" else 0) or seq.shape[0] == b.n
        for i in range(b.n):
            assert seq[i].tobytes() == want_seq[i], f"read line {i}"
            assert q[i].tobytes() == want_q[i], f"quality line {i}"


def test_oracle_parse_and_quality_stats_match_reference(tmp_path, oracle_lib):
    """orc_parse_fastq (the checker of scb_submit_fastq) against the UNMODIFIED reference: the harness feeds the same reads to the
    reference's output_quality with arithmetic coding on (_no_ac = 0) in a fresh process (its `prev` is a function-static) and its
    ac_freq3 / ac_freq4 globals must equal the oracle's input-order statistics; names and payload must equal what the SoA holds."""
    import subprocess
    import sys
    harness = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_harness.so")
    if not os.path.exists(harness):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cores = make_cores(941, [(8, 64), (9, 32)])
    b = synth.make_batch(3000, 70, seed=941, n_frac=0.02, high_entropy=True)
    b2 = synth.make_batch(2000, 70, seed=942, n_frac=0.0)
    d = str(tmp_path)
    synth.write_fastq(b, d + "/a.fastq")
    synth.write_fastq(b2, d + "/b.fastq")
    ta, tb = open(d + "/a.fastq", "rb").read(), open(d + "/b.fastq", "rb").read()
    # oracle: two texts one after the other (statistics and prev carried across)
    st = orc.new_quality_stats()
    seq_a, q_a, nm_a, off_a = orc.parse_fastq(ta, 70, 33, st)
    seq_b, q_b, nm_b, off_b = orc.parse_fastq(tb[:-1], 70, 33, st)          # last line without its newline
    assert np.array_equal(seq_a, b.seq) and np.array_equal(seq_b, b2.seq)
    assert np.array_equal(q_a, orc.quality_payload(b.qual, b.seq, 33)) and np.array_equal(q_b, orc.quality_payload(b2.qual, b2.seq, 33))
    assert nm_a.tobytes() == b.names.tobytes() and np.array_equal(off_a, b.name_off) and nm_b.tobytes() == b2.names.tobytes()
    np.savez(d + "/in.npz", seq=np.concatenate([b.seq, b2.seq]), qual=np.concatenate([b.qual, b2.qual]),
             names=np.concatenate([b.names, b2.names]), off=np.concatenate([b.name_off, b2.name_off[1:] + b.name_off[-1]]))
    open(d + "/cores.txt", "w").write("\n".join(cores) + "\n")
    code = f"""
import ctypes as C, numpy as np
H = C.CDLL({harness!r})
H.refh_init.restype = C.c_double; H.refh_init.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
H.refh_run.restype = C.c_double; H.refh_run.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
H.refh_set_no_ac.argtypes = [C.c_int]; H.refh_get_stats.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
z = np.load({d + '/in.npz'!r})
H.refh_set_no_ac(0)
H.refh_init({d + '/cores.txt'!r}.encode(), 70, 0, 0, 1, 4 << 30)
p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
seq, qual, names, off = [np.ascontiguousarray(z[k]) for k in ('seq', 'qual', 'names', 'off')]
nch = C.c_int()
H.refh_run(seq.shape[0], p(seq), p(qual), p(names), p(off), None, None, 33, {d!r}.encode(), C.byref(nch), None, None)
f3 = np.zeros(6400, dtype=np.uint64); f4 = np.zeros(512000, dtype=np.uint64)
H.refh_get_stats(0, p(f3), p(f4))
np.savez({d + '/ref_stats.npz'!r}, f3=f3, f4=f4)
"""
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    ref = np.load(d + "/ref_stats.npz")
    assert np.array_equal(ref["f3"], st["freq3"]), "ac_freq3 differs from the reference's output_quality"
    assert np.array_equal(ref["f4"], st["freq4"]), "ac_freq4 differs from the reference's output_quality"
    assert int(st["freq3"].sum()) == 5000 * 70 - 1 and int(st["freq4"].sum()) == 512000 + 5000 * 70 - 2
    with pytest.raises(ValueError):
        orc.parse_fastq(ta[:len(ta) // 2 + 7], 70, 33)

#!/usr/bin/env python3
"""Builds oracle/_ref/scalce_scb: the reference CLI with ONLY the boosting transform replaced by libscalce_b200.so.

A copy of /root/reference/compress.cpp gets five textual edits (the call sites INTEGRATION.md lists) and an #include of
oracle/dropin/scb_glue.h; every other reference source is compiled unmodified (the objects oracle/Makefile already
builds). The patched copy lives in oracle/_ref/dropin/ (git-ignored) - reference sources never enter the repository.
TEST INFRASTRUCTURE (tests/test_gpu_dropin.py). Needs /root/reference; the built binary travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
ROOT = os.path.dirname(ORACLE)
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(ORACLE, "_ref")
BIN = os.path.join(OUT, "scalce_scb")


def patch(src: str) -> str:
    def sub(old, new, count=1):
        nonlocal src
        assert src.count(old) >= 1, f"anchor not found in compress.cpp: {old[:60]!r}"
        src = src.replace(old, new, count)

    # 1. the glue, in front of thread() (all globals it needs are declared above)
    sub("void *thread(void *vt) {", '#include "scb_glue.h"\n\nvoid *thread(void *vt) {')
    # 2. compress.cpp:673-715: search / pack / bucket insert / size accounting / flush trigger -> append to the SoA batch
    a = src.index("    int n = aho_search(read, trie, &bucket);")
    tail = "        total_size = 0;\n      }\n      pthread_spin_unlock(&w_spin);\n    }\n"
    b = src.index(tail, a) + len(tail)
    src = src[:a] + "    scb_glue_append(name, read, qual, read2, qual2, qmap);\n" + src[b:]
    # 3. compress.cpp:732-733: core-set load -> after get_quality_stats, when the read lengths are known
    sub("  trie =\n      pattern_path[0] ? read_patterns_from_file(pattern_path) : read_patterns();", "  trie = 0;")
    sub("  get_quality_stats(input, files[0], qmap);", "  get_quality_stats(input, files[0], qmap);\n  scb_glue_create(pattern_path);")
    # 4. compress.cpp:799-801: final dump_trie -> one flush of everything submitted
    sub("  if (total_size) {\n    dump_trie(temp_file_count++, trie);\n  }", "  scb_glue_flush();")
    # 5. compress.cpp:821, 834
    sub("  aho_trie_free(trie);", "  scb_glue_destroy();")
    src = src.replace("unbuck()", "scb_glue_unbucketed()")
    return src


def build(verbose=False):
    if not os.path.isdir(REF):
        raise SystemExit("make_dropin.py: needs the reference sources (" + REF + ")")
    subprocess.check_call(["make", "-s", "-C", ORACLE, "ref"])          # unmodified objects + _ref/HELP.o + _ref/patterns.o
    d = os.path.join(OUT, "dropin")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(REF, "compress.cpp")) as f:
        src = patch(f.read())
    with open(os.path.join(d, "compress_scb.cpp"), "w") as f:
        f.write(src)
    lib_dir = os.path.join(ROOT, "scalce_b200")
    flags = ["-O3", "-DNDEBUG", "-w", "-I" + os.path.join(ORACLE, "shim"), "-I" + REF, "-I" + HERE, "-I" + os.path.join(ROOT, "include"),
             "-D_FILE_OFFSET_BITS=64", "-D_LARGEFILE64_SOURCE", '-DSCALCE_VERSION="2.8"']
    subprocess.check_call(["g++", "-c", *flags, os.path.join(d, "compress_scb.cpp"), "-o", os.path.join(d, "compress_scb.o")])
    objs = [os.path.join(OUT, x + ".o") for x in ("const", "buffio", "arithmetic", "main", "names", "qualities", "reads", "decompress", "HELP", "patterns")]
    subprocess.check_call(["g++", os.path.join(d, "compress_scb.o"), *objs, "-L" + lib_dir, "-lscalce_b200", "-Wl,-rpath,$ORIGIN/../../scalce_b200",
                           "-lm", "-lpthread", "-lz", "/usr/lib/x86_64-linux-gnu/libbz2.so.1.0", "-o", BIN])
    if verbose:
        print(BIN)
    return BIN


if __name__ == "__main__":
    build(verbose=True)

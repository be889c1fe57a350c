"""Helper for tests/test_host_cpu.py: runs bench.py's native arm with the device and the library replaced by
fakes (no GPU, no kernels), so that the control flow and the JSON line of the bench are exercised on the CPU.
Usage: python tests/bench_dry_run.py <bench.py arguments...>   -> prints the bench's JSON line."""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self: self
torch.cuda.mem_get_info = lambda *a, **k: (150 << 30, 180 << 30)


class _Ev:
    def __init__(self, **k):
        pass

    def record(self):
        pass

    def elapsed_time(self, o):
        return 1.0


torch.cuda.Event = _Ev

from scalce_b200 import binding, synth   # noqa: E402


def _make_batch_cuda(N, L, seed=1, device=None, paired=False, L2=None, high_entropy=False, **k):
    b = synth.make_batch(N, L, seed=seed, paired=paired, L2=L2, high_entropy=high_entropy)
    names = np.frombuffer(b"".join(b"SYN.%09d" % i for i in range(N)), dtype=np.uint8).copy()
    d = dict(seq=torch.from_numpy(b.seq), qual=torch.from_numpy(b.qual), names=torch.from_numpy(names),
             name_off=torch.arange(N + 1, dtype=torch.int64) * 13)
    if paired:
        d["seq2"], d["qual2"] = torch.from_numpy(b.seq2), torch.from_numpy(b.qual2)
    return d


synth.make_batch_cuda = _make_batch_cuda


class _Res:
    def __init__(self, n, paired=False):
        self.device_ms, self.n_chunks, self.n_reads = 2.0, 2, n
        self.chunk_off = [[0, 10, 20], [0, 30, 60], [0, 100, 200], [0, 8, 16], [0, 0, 0], [0, 0, 0]]
        if paired:
            self.chunk_off[4], self.chunk_off[5] = [0, 40, 80], [0, 100, 200]


class _FakeTransform:
    STAGES = binding.BoostTransform.STAGES
    resolve_rounds = 7
    kernel_launches = 90
    resolve_engine = 0
    device_bytes = 123456

    def __init__(self, cores, L, L2=0, paired=False, device=0, emit_merged=True):
        self._h, self.n, self.paired = 1, 0, paired

    def table_info(self):
        return dict(n_cores=2048, n_states=9000, n_buckets=2048, smem_resident=True)

    def reset_counts(self):
        pass

    def submit_device(self, n, *a):
        self.n = n

    def submit(self, seq, *a):
        self.n = seq.shape[0]

    def flush(self):
        return _Res(self.n, self.paired)

    flush_closed = flush

    def stage_ms(self):
        return dict(zip(self.STAGES, [1.0, 2.0, 0.1, 0.5, 0.1, 3.0, 0.0, 0.05]))

    def close(self):
        pass


class _FakeLib:
    k = 0

    def scb_kernel_launches(self, h):
        self.k += 45
        return self.k

    def scb_copy_stream(self, h, k, c, dst, nbytes):
        import ctypes
        ctypes.memset(dst, 0x5a, int(nbytes))
        return 0


binding.BoostTransform = _FakeTransform
_lib = _FakeLib()
binding.load_library = lambda path=None: _lib

import bench   # noqa: E402

sys.argv = ["bench.py"] + sys.argv[1:]
out = io.StringIO()
with contextlib.redirect_stdout(out):
    bench.main()
print([ln for ln in out.getvalue().splitlines() if ln.startswith("{")][-1])

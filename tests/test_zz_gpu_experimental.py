"""Runs tests/experimental_cases.py (parity of the opt-in variants that were written without GPU access) in one child
process, last in the suite, with a hard time limit. Outcome: pass if every case passes, xfail otherwise (with the
child's summary) - the variants are off by default, so their state must not decide whether the product suite is green."""
import os
import subprocess
import sys

import pytest

# Off unless SCB_RUN_EXPERIMENTAL=1: the variants have never run on a B200, and the round-end suite shares its box with the
# smoke run and the bench - not the place for a first run of unverified kernels. tools/round2_ab.sh runs the cases directly.
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("SCB_RUN_EXPERIMENTAL", "0") in ("", "0"),
                                                   reason="experimental variants: set SCB_RUN_EXPERIMENTAL=1 (or run tools/round2_ab.sh)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_experimental_variants_in_child_process():
    env = dict(os.environ, SCB_TEST_EXPERIMENTAL="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "experimental_cases.py"), "-q", "-rfEs", "-p", "no:cacheprovider"]
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    except subprocess.TimeoutExpired as ex:
        tail = (ex.stdout or b"").decode(errors="replace")[-1500:]
        pytest.xfail("experimental variants: child exceeded 300 s\n" + tail)
    out = r.stdout.decode(errors="replace")
    print(out[-4000:])
    if r.returncode != 0:
        pytest.xfail("experimental variants (off by default) are not all green yet:\n" + out[-2500:])

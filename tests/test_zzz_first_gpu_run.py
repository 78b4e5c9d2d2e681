"""Isolated runner for the GPU parity cases of the NEWEST kernels (tests/_first_gpu_run_cases.py):
each case runs HERE in its own pytest process with a time limit, so that a hang or a fault of a
young kernel cannot take the other tests (or this pytest process, or the CUDA context its session
fixture holds) with it.  The file runs LAST (name).  No xfail: a mismatch, crash or timeout FAILS,
and a case that skipped in its child process is reported as SKIPPED here (round 1 counted such
skips as passes).  Cases move to tests/test_gpu_parity_variants.py once they are old."""
import os
import signal
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES_FILE = os.path.join(ROOT, "tests", "_first_gpu_run_cases.py")

pytestmark = [pytest.mark.gpu]

CASE_TIMEOUT_S = 240     # one case; the slowest (1.2 Gbp packed store, 2^30-base poly-A) take well under a minute on a B200
FILE_BUDGET_S = 900      # all cases together: a systematic hang must not eat the caller's time limit
_state = {"t0": None, "gpu_lost": False}

CASES = [
    "test_partition_wide2",
    "test_fingerprint_self_checks",
    "test_sparse_radix_rounds",
]


def _run(cmd, timeout):
    """run cmd in its own process group; on timeout the whole group is killed (torchrun grandchildren too)"""
    p = subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         stdin=subprocess.DEVNULL, start_new_session=True)
    try:
        out, _ = p.communicate(timeout=timeout)
        return p.returncode, out
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        try:
            out, _ = p.communicate(timeout=30)
        except subprocess.TimeoutExpired:
            out = ""
        return None, out or ""


def _gpu_answers():
    rc, _ = _run([sys.executable, "-c", "import torch; torch.zeros(1, device='cuda:0').item(); torch.cuda.synchronize()"], 120)
    return rc == 0


def case_names():
    """the test functions of the cases file, by a text scan (no import: the static check below must hold on CPU)"""
    names = []
    with open(CASES_FILE) as f:
        for ln in f:
            if ln.startswith("def test_"):
                names.append(ln[4:ln.index("(")])
    return names


@pytest.mark.parametrize("case", CASES)
def test_first_run(case):
    if _state["t0"] is None:
        _state["t0"] = time.monotonic()
    if _state["gpu_lost"]:
        pytest.fail("skipped: the GPU stopped answering after an earlier case")
    left = FILE_BUDGET_S - (time.monotonic() - _state["t0"])
    if left < 30:
        pytest.fail("skipped: this file's %d s budget is spent" % FILE_BUDGET_S)
    cmd = [sys.executable, "-m", "pytest", CASES_FILE, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", case]
    rc, out = _run(cmd, min(CASE_TIMEOUT_S, left))
    if rc is None:
        _state["gpu_lost"] = not _gpu_answers()
        pytest.fail("%s: no result within the time limit (killed)%s\n%s"
                    % (case, "; the GPU no longer answers" if _state["gpu_lost"] else "", out[-3000:]))
    assert rc == 0, "%s: exit code %d\n%s" % (case, rc, out[-6000:])
    if " passed" not in out:
        if " skipped" in out:
            pytest.skip("skipped: the case skipped in its own process\n" + out[-600:])
        pytest.fail("%s: neither passed nor skipped\n%s" % (case, out[-2000:]))

"""GPU parity cases of this round's NEW kernels.  NOT collected directly (file name):
tests/test_zzz_first_gpu_run.py runs every case in its own process with a time limit, so that a
hang or a fault of a young kernel cannot take the other tests (or the pytest process) with it.
KC_FIRST_RUN_SCALE < 1 shrinks every input so that the same Python runs on the CPU emulator
(tests/emu/run_under_shim.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu]

SCALE = float(os.environ.get("KC_FIRST_RUN_SCALE", "1"))
FULL = SCALE == 1.0


def sz(n):
    return n if FULL else max(2000, int(n * SCALE))


def to_dev(arr):
    import torch
    a = np.ascontiguousarray(arr, dtype=np.uint8)
    buf = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda:0")
    if a.size:
        buf[: a.size] = torch.from_numpy(a.copy())
    return buf


def _dense(ctx, kmerlib, data, k, algo):
    import torch
    table = torch.zeros(kmerlib.num_kmers(k), dtype=torch.int32, device="cuda:0")
    d = to_dev(data)
    ctx.count_dense_range(d, data.size, 0, data.size, k, table, algo=algo)
    torch.cuda.synchronize()
    return table.cpu().numpy().view(np.uint32)

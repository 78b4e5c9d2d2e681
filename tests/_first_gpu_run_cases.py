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


def test_partition_wide2(ctx, kmerlib, oracle):
    """KC_DENSE_PARTITION_WIDE2 (k = 12; 16 records per lane and super-step, one shared atomic per record,
    warp-cooperative bin flush) against the oracle (genome with N runs, dirty bytes, unaligned pointer),
    against the five-sub-table partition path at 1 Gbp, and on 2^26 'A's (every record into one bin:
    retries, region overflow, nibble wraps -> 32-bit recount)"""
    import torch
    n = sz(40_000_000)
    genome = oracle.gen_genome(0xB2000003, n, 40, 400, 12, 0, n)
    want, _ = oracle.count_dense(genome, 12)
    assert (_dense(ctx, kmerlib, genome, 12, kmerlib.DENSE_PARTITION_WIDE2) == want).all()
    rng = np.random.default_rng(11)
    alpha = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTNacgt\n\0|>", dtype=np.uint8)
    dirty = alpha[rng.integers(0, alpha.size, sz(30_000_000))]
    want, _ = oracle.count_dense(dirty, 12)
    table = torch.zeros(kmerlib.num_kmers(12), dtype=torch.int32, device="cuda:0")
    d = to_dev(np.concatenate([np.zeros(5, np.uint8), dirty]))
    ctx.count_dense_range(d[5:], dirty.size, 0, dirty.size, 12, table, algo=kmerlib.DENSE_PARTITION_WIDE2)
    torch.cuda.synchronize()
    assert (table.cpu().numpy().view(np.uint32) == want).all()
    L = sz(1 << 30)
    data = ctx.gen_genome(0xB2000003, L, 300, 3000, 12, 0, L)
    a = torch.zeros(kmerlib.num_kmers(12), dtype=torch.int32, device="cuda:0")
    b = torch.zeros_like(a)
    ctx.count_dense_range(data, L, 0, L, 12, a, algo=kmerlib.DENSE_PARTITION)
    ctx.count_dense_range(data, L, 0, L, 12, b, algo=kmerlib.DENSE_PARTITION_WIDE2)
    torch.cuda.synchronize()
    assert bool((a == b).all())
    del data, a, b
    P = sz(1 << 26)
    poly = torch.full((P,), ord("A"), dtype=torch.uint8, device="cuda:0")
    t = torch.zeros(kmerlib.num_kmers(12), dtype=torch.int32, device="cuda:0")
    ctx.count_dense_range(poly, P, 0, P, 12, t, algo=kmerlib.DENSE_PARTITION_WIDE2)
    torch.cuda.synchronize()
    assert int(t[0].item()) == P - 11 and int(t.to(torch.int64).sum().item()) == P - 11


def test_fingerprint_self_checks(ctx, kmerlib, oracle):
    """csrc/check.cu: the window fingerprint of an input (one streaming scan) equals the numpy value from the oracle's
    count, and the fingerprint of every GPU count of it (sparse auto / hash / radix, dense); at 20 M windows, where the
    oracle does not go, input and result fingerprints agree and the keys are strictly ascending"""
    import torch
    from kmerb200.distributed import mix64_np
    reads = oracle.gen_reads(0xB2000004, 3_000_000, 150, 200, 0, sz(30_000) if FULL else 600)
    d = to_dev(reads)
    for k in (12, 21, 31):
        wk, wc, _ = oracle.count_sparse(reads, k)
        with np.errstate(over="ignore"):
            want = (int((mix64_np(wk) * wc.astype(np.uint64)).sum(dtype=np.uint64)), int(wc.astype(np.uint64).sum()))
        assert ctx.window_fingerprint(d, reads.size, k) == want, k
        for algo in (kmerlib.SPARSE_AUTO, kmerlib.SPARSE_HASH, kmerlib.SPARSE_RADIX):
            sp = ctx.count_sparse(d, reads.size, k, algo)
            assert ctx.sparse_fingerprint(sp) == want + (0,), (k, algo)
            sp.close()
        if k == 12:
            t = torch.zeros(kmerlib.num_kmers(k), dtype=torch.int32, device="cuda:0")
            ctx.count_dense_range(d, reads.size, 0, reads.size, k, t)
            torch.cuda.synchronize()
            assert ctx.dense_fingerprint(t, k) == want
    nreads = 150_000 if FULL else 700
    big = ctx.gen_reads(0xB2000004, 50_000_000, 150, 200, 0, nreads)
    fin = ctx.window_fingerprint(big, nreads * 151, 21)
    assert fin[1] == nreads * 130
    sp = ctx.count_sparse(big, nreads * 151, 21, kmerlib.SPARSE_AUTO)
    assert ctx.sparse_fingerprint(sp) == fin + (0,)
    sp.close()


def test_sparse_radix_rounds(kmerlib):
    """KC_SPARSE_RADIX in several ROUNDS (what inputs too large for one pass take: kc_sparse_radix_plan picks round_bits
    from leaf capacity and device memory; KC_SPARSE_RADIX_RBITS forces it at test sizes, read once per process, hence
    the child): the result must equal the hash path's and the input's window fingerprint, keys strictly ascending"""
    code = r"""
import os, sys
sys.path.insert(0, os.path.join(%r, "dna-kmeres-parallel_b200"))
import numpy as np, kmerb200 as K
ctx = K.Context(0)
for k, nreads in ((21, %d), (31, %d), (13, %d)):
    reads = ctx.gen_reads(0xB2000004, 50_000_000, 150, 200, 0, nreads)
    nb = nreads * 151
    a = ctx.count_sparse(reads, nb, k, K.SPARSE_HASH)
    b = ctx.count_sparse(reads, nb, k, K.SPARSE_RADIX | K.SPARSE_NO_FALLBACK)
    ka, ca = a.to_host(); kb, cb = b.to_host()
    assert len(a) == len(b) and (ka == kb).all() and (ca == cb).all(), k
    assert ctx.sparse_fingerprint(b) == ctx.window_fingerprint(reads, nb, k) + (0,), k
print("ROUNDS_OK")
""" % (ROOT, sz(150_000) if FULL else 700, sz(100_000) if FULL else 500, sz(100_000) if FULL else 500)
    import subprocess
    # KC_SPARSE_RADIX_RUNLIST: a first temporary run list this short, so that the count is repeated with the exact size
    # the leaf kernel reported (run_count's second try), with one round and with appended rounds
    for extra in (dict(KC_SPARSE_RADIX_RBITS="1"), dict(KC_SPARSE_RADIX_RBITS="3"), dict(KC_SPARSE_RADIX_RBITS="6"),
                  dict(KC_SPARSE_RADIX_RBITS="0", KC_SPARSE_RADIX_RUNLIST="5000"),
                  dict(KC_SPARSE_RADIX_RBITS="2", KC_SPARSE_RADIX_RUNLIST="5000")):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=200, env=dict(os.environ, **extra))
        assert r.returncode == 0 and "ROUNDS_OK" in r.stdout, str(extra) + r.stdout[-1500:] + r.stderr[-3000:]


"""The N>1 host path (shards, halo, reduce, owner all-to-all) at world_size 2 on the
CPU with gloo; the oracle stands in for the per-rank GPU engine."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29653", os.path.join(ROOT, "tests", "_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_WORKER_OK world=2" in r.stdout


def test_world2_gloo_radix():
    """range-sharded sparse radix path: production host logic over gloo, emulator kernels per rank"""
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29657", os.path.join(ROOT, "tests", "_gloo_radix_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", KC_SPARSE_RADIX_SHAPE="small", KC_EMU_SMS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_RADIX_WORKER_OK world=2" in r.stdout


def test_nccl_radix_worker_dry_run():
    """the GPU worker of the range-sharded radix path itself (tests/_nccl_radix_worker.py: kmerb200.Context's
    radix_plan / scatter / count methods, count_sparse_sharded_gpu's dispatch and overflow vote), unmodified,
    on the CPU: emulator library + torch CUDA stand-ins + gloo (tests/emu/run_under_shim.py), small cases"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29658", os.path.join(ROOT, "tests", "emu", "run_under_shim.py"),
           os.path.join(ROOT, "tests", "_nccl_radix_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", KC_SPARSE_RADIX_SHAPE="small", KC_EMU_SMS="4", KC_NCCL_RADIX_CASES="21:600,31:500")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "NCCL_RADIX_WORKER_OK world=2" in r.stdout


def test_nccl_perseq_worker_dry_run():
    """SURVEY §8e, per-sequence mode: the GPU worker (tests/_nccl_perseq_worker.py: sequences sharded at sequence
    starts, kc_count_per_seq per rank, all-gather of the columns, the distance step on the gathered table), unmodified,
    on the CPU with two ranks: emulator library + torch CUDA stand-ins + gloo"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29659", os.path.join(ROOT, "tests", "emu", "run_under_shim.py"),
           os.path.join(ROOT, "tests", "_nccl_perseq_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", KC_EMU_SMS="4", KC_NCCL_PERSEQ_CASES="23:3000:3,7:2000:5,2:500:4,1:300:2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "NCCL_PERSEQ_WORKER_OK world=2" in r.stdout


def test_shard_seqs_cover(kmerlib):
    """sequence shards: contiguous, disjoint, cover all sequences, cut at sequence starts, near-equal bytes"""
    rng = np.random.default_rng(4)
    for n in (0, 1, 2, 7, 100, 1000):
        lens = rng.integers(1, 5000, n)
        offs = np.cumsum(np.concatenate([[0], lens + 1])).astype(np.int64)
        for world in (1, 2, 3, 8):
            prev = 0
            sizes = []
            for r in range(world):
                s0, s1 = kmerlib.shard_seqs(offs, r, world)
                assert s0 == prev and s1 >= s0
                prev = s1
                sizes.append(int(offs[s1] - offs[s0]) if n else 0)
            assert prev == n
            if n >= 100:   # no shard more than one sequence away from its fair share
                assert max(sizes) <= offs[-1] / world + 5001 * 2, (n, world, sizes)


def test_mix64_np_matches_engine(kmerlib, oracle):
    from kmerb200 import distributed as D
    xs = np.array([0, 1, 0xDEADBEEF, (1 << 62) - 1, 12345678901234567], dtype=np.uint64)
    got = D.mix64_np(xs)
    assert [int(g) for g in got] == [kmerlib.mix64(int(x)) for x in xs] == [oracle.mix64(int(x)) for x in xs]


def test_merge_and_bucket_np():
    from kmerb200 import distributed as D
    keys = np.array([5, 9, 5, 1, 9, 9], dtype=np.uint64)
    counts = np.array([1, 2, 3, 4, 5, 6], dtype=np.uint32)
    k, c = D.merge_np(keys, counts)
    assert k.tolist() == [1, 5, 9] and c.tolist() == [4, 4, 13]
    bk, bc, sizes = D.bucket_by_owner_np(keys, counts, 3)
    assert sizes.sum() == 6 and sorted(bk.tolist()) == sorted(keys.tolist())
    start = 0
    for o in range(3):
        assert ((D.mix64_np(bk[start:start + sizes[o]]) % np.uint64(3)) == o).all()
        start += sizes[o]

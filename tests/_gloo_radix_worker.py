"""world_size-2 worker for tests/test_sharding_gloo.py::test_world2_gloo_radix: the production
host logic of the range-sharded sparse radix path (kmerb200.distributed.count_sparse_radix_sharded:
plan agreement, overflow vote, equal-split all-to-all of the partition-major slabs, per-rank
count) over gloo, with the emulator build of the kernels standing in for the GPU on each rank."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200"), os.path.join(ROOT, "tests", "emu")):
    sys.path.insert(0, p)
import oracle as O  # noqa: E402
import emu_harness as H  # noqa: E402
from kmerb200 import distributed as D  # noqa: E402


class Full(RuntimeError):
    pass


class EmuEngine:
    """radix_plan / radix_scatter / radix_count on torch CPU tensors, computed by libkmerb200_emu.so"""

    def __init__(self):
        self.ctx = H.EmuContext()

    def _check(self, rc):
        if rc == -5:
            msg = self.ctx.L.kc_last_error(self.ctx.h).decode()
            sys.stderr.write("[rank %d] %s\n" % (dist.get_rank(), msg))
            raise Full(msg)
        self.ctx.check(rc)

    def radix_plan(self, max_windows, k, world, min_round_bits=0):
        plan = H.RadixPlan()
        self.ctx.L.kc_sparse_radix_plan_rounds.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
        self._check(self.ctx.L.kc_sparse_radix_plan_rounds(self.ctx.h, max_windows, k, world, min_round_bits, C.byref(plan)))
        return plan

    def radix_scatter(self, reads, nbytes, plan, rnd=0, out=None):
        base, p = self.ctx.upload(reads.numpy(), 1)
        d_s, d_c = self.ctx.alloc(plan.slab_bytes), self.ctx.alloc(plan.counts_bytes)
        try:
            self._check(self.ctx.L.kc_sparse_radix_scatter_round(self.ctx.h, p, nbytes, C.byref(plan), rnd, d_s, d_c))
            return (torch.from_numpy(self.ctx.download(d_s, plan.slab_bytes, np.uint8)),
                    torch.from_numpy(self.ctx.download(d_c, plan.counts_bytes, np.int32)))
        finally:
            for q in (base, d_s, d_c):
                self.ctx.free(q)

    def radix_count(self, plan, slabs, counts, nsrc, part_first, nparts, rnd=0):
        b1, d_s = self.ctx.upload(slabs.numpy())
        b2, d_c = self.ctx.upload(counts.numpy())
        sp = C.c_void_p()
        try:
            self._check(self.ctx.L.kc_sparse_radix_count_round(self.ctx.h, C.byref(plan), rnd, d_s, d_c, nsrc, part_first, nparts, C.byref(sp)))
        finally:
            self.ctx.free(b1)
            self.ctx.free(b2)
        n = int(self.ctx.L.kc_sparse_size(sp))
        keys, cnts = np.empty(n, np.uint64), np.empty(n, np.uint32)
        self.ctx.check(self.ctx.L.kc_sparse_copy_to_host(self.ctx.h, sp, keys.ctypes.data, cnts.ctypes.data))
        self.ctx.L.kc_sparse_free(sp)
        return keys, cnts

    def radix_count_append(self, plan, slabs, counts, nsrc, part_first, nparts, rnd, acc=None):
        """kc_sparse_radix_count_round_append: `acc` = the kc_sparse handle that grows round by round"""
        b1, d_s = self.ctx.upload(slabs.numpy())
        b2, d_c = self.ctx.upload(counts.numpy())
        h = acc if acc is not None else C.c_void_p()
        try:
            self._check(self.ctx.L.kc_sparse_radix_count_round_append(self.ctx.h, C.byref(plan), rnd, d_s, d_c, nsrc, part_first, nparts,
                                                                      C.byref(h)))
        finally:
            self.ctx.free(b1)
            self.ctx.free(b2)
        return h

    def finish(self, acc):  # handle -> (keys, counts) on the host
        n = int(self.ctx.L.kc_sparse_size(acc))
        keys, cnts = np.empty(n, np.uint64), np.empty(n, np.uint32)
        self.ctx.check(self.ctx.L.kc_sparse_copy_to_host(self.ctx.h, acc, keys.ctypes.data, cnts.ctypes.data))
        self.ctx.L.kc_sparse_free(acc)
        return keys, cnts

    def sparse_concat(self, parts):  # the rounds' pieces of this rank, ascending
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = EmuEngine()
    for k, nreads, genome in ((21, 700, 80_000), (31, 500, 60_000)):
        r0, r1 = D.shard_reads(nreads, rank, world)
        mine = O.gen_reads(0xB2000004, genome, 100, 50, r0, r1 - r0)
        res = D.count_sparse_radix_sharded(eng, torch.from_numpy(mine.copy()), mine.size, k, table_full=(Full,))
        assert res is not None, "uniform reads must not overflow"
        keys, cnts = res
        gathered = [None] * world
        dist.all_gather_object(gathered, (keys, cnts))
        if rank == 0:
            allk = np.concatenate([g[0] for g in gathered])  # one round: rank order = code order, no sort here
            allc = np.concatenate([g[1] for g in gathered])
            if int(os.environ.get("KC_SPARSE_RADIX_RBITS", "0")) > 0:  # several rounds: every rank holds one ascending range per round
                for g in gathered:
                    assert (np.diff(g[0].astype(np.int64)) > 0).all(), "a rank's keys must be ascending"
                order = np.argsort(allk, kind="stable")
                allk, allc = allk[order], allc[order]
            whole = O.gen_reads(0xB2000004, genome, 100, 50, 0, nreads)
            wk, wc, _ = O.count_sparse(whole, k)
            assert allk.size == wk.size and (allk == wk).all() and (allc == wc).all(), "sharded radix != whole (k=%d)" % k
    # one rank's input overflows a leaf (poly-A): EVERY rank must report None, so all fall back together
    data = np.full(40_000, ord("A"), dtype=np.uint8) if rank == 1 else O.gen_reads(5, 6000, 100, 50, 0, 300)
    res = D.count_sparse_radix_sharded(eng, torch.from_numpy(data.copy()), data.size, 21, table_full=(Full,))
    assert res is None, "overflow on rank 1 must make every rank fall back"
    dist.barrier()
    if rank == 0:
        print("GLOO_RADIX_WORKER_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

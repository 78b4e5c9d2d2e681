"""world_size-2 worker for tests/test_sharding_gloo.py (launched with torch.distributed.run).
The oracle stands in for the GPU engine; everything else is the production host logic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200")):
    sys.path.insert(0, p)
import oracle as O  # noqa: E402
from kmerb200 import distributed as D  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # ---------------- dense: window shards + halo, reduce ----------------------
    L, k, seed = 400_003, 12, 0xB2000003
    for k in (3, 12):
        full = O.gen_genome(seed, L, 5, 200, 12, 0, L) if rank == 0 else None
        table = torch.zeros(4 ** k, dtype=torch.int32)

        def make_shard(bb, be):
            return O.gen_genome(seed, L, 5, 200, 12, bb, be - bb)  # each rank generates only its bytes

        def count_range(shard, n, wb, we, t):
            tmp, _ = O.count_dense_range(shard, k, wb, we)
            t += torch.from_numpy(tmp.view(np.int32))

        D.count_dense_sharded(count_range, make_shard, L, k, table, rank, world, dst=0)
        if rank == 0:
            want, _ = O.count_dense(full, k)
            assert (table.numpy().view(np.uint32) == want).all(), "dense k=%d sharded != whole" % k
    # ---------------- sparse: read shards, owner buckets, all-to-all ------------
    nreads, k = 4000, 21
    r0, r1 = D.shard_reads(nreads, rank, world)
    mine = O.gen_reads(0xB2000004, 60_000, 150, 200, r0, r1 - r0)
    keys, counts, _ = O.count_sparse(mine, k)
    bk, bc, sizes = D.bucket_by_owner_np(keys, counts, world)
    rk, rc = D.exchange_by_owner(torch.from_numpy(bk.view(np.int64)), torch.from_numpy(bc.view(np.int32)), sizes)
    ok, oc = D.merge_np(rk.numpy().view(np.uint64), rc.numpy().view(np.uint32))
    assert ((D.mix64_np(ok) % np.uint64(world)) == rank).all()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, oc))
    if rank == 0:
        allk = np.concatenate([g[0] for g in gathered])
        allc = np.concatenate([g[1] for g in gathered])
        order = np.argsort(allk)
        whole = O.gen_reads(0xB2000004, 60_000, 150, 200, 0, nreads)
        wk, wc, _ = O.count_sparse(whole, k)
        assert (allk[order] == wk).all() and (allc[order] == wc).all(), "sparse sharded != whole"
        print("GLOO_WORKER_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""world_size-N GPU worker for tests/test_multi_gpu.py (torch.distributed.run, NCCL):
dense window shards + reduce, sparse read shards + owner all-to-all, both against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200")):
    sys.path.insert(0, p)
import oracle as O  # noqa: E402
import kmerb200  # noqa: E402
from kmerb200 import distributed as D  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = kmerb200.Context(local)
    # dense, k=12 (partition path on each shard) and k=5
    L, seed = 200_000_003, 0xB2000003
    for k in (12, 5):
        table = torch.zeros(4 ** k, dtype=torch.int32, device=dev)

        def make_shard(bb, be):
            return ctx.gen_genome(seed, L, 50, 2000, 12, bb, be - bb)

        def count_range(shard, n, wb, we, t):
            ctx.count_dense_range(shard, n, wb, we, k, t)

        D.count_dense_sharded(count_range, make_shard, L, k, table, rank, world, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            whole = O.gen_genome(seed, L, 50, 2000, 12, 0, L)
            want, _ = O.count_dense(whole, k, threads=min(os.cpu_count() or 4, 16))
            assert (table.cpu().numpy().view(np.uint32) == want).all(), "dense k=%d sharded != oracle" % k
    # sparse, k=21, hash-sharded all-to-all
    nreads, k = 200_000, 21
    r0, r1 = D.shard_reads(nreads, rank, world)
    reads = ctx.gen_reads(0xB2000004, 2_000_000, 150, 200, r0, r1 - r0)
    sp = D.count_sparse_sharded_gpu(ctx, reads, (r1 - r0) * 151, k, kmerb200.SPARSE_HASH)
    keys, counts = sp.to_host()
    assert ((D.mix64_np(keys) % np.uint64(world)) == rank).all()
    gathered = [None] * world
    dist.all_gather_object(gathered, (keys, counts))
    if rank == 0:
        allk = np.concatenate([g[0] for g in gathered])
        allc = np.concatenate([g[1] for g in gathered])
        order = np.argsort(allk)
        whole = O.gen_reads(0xB2000004, 2_000_000, 150, 200, 0, nreads)
        wk, wc, _ = O.count_sparse(whole, k)
        assert (allk[order] == wk).all() and (allc[order] == wc).all(), "sparse sharded != oracle"
        print("NCCL_WORKER_OK world=%d" % world)
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

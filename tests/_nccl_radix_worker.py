"""world_size-N GPU worker (torch.distributed.run, NCCL) for the range-sharded sparse radix path:
scatter on every rank, one equal-split all-to-all of the partition-major slabs, count per rank;
the ranks' results concatenated in rank order must be the oracle's sorted whole."""
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
    # KC_NCCL_RADIX_CASES="21:600,31:500": smaller cases for the CPU dry run of this file (tests/test_sharding_gloo.py)
    cases = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("KC_NCCL_RADIX_CASES", "21:200000,31:100000").split(",")]
    for k, nreads in cases:
        r0, r1 = D.shard_reads(nreads, rank, world)
        reads = ctx.gen_reads(0xB2000004, 20_000_000, 150, 200, r0, r1 - r0)
        sp = D.count_sparse_sharded_gpu(ctx, reads, (r1 - r0) * 151, k, kmerb200.SPARSE_RADIX | kmerb200.SPARSE_NO_FALLBACK)
        keys, counts = sp.to_host()
        gathered = [None] * world
        dist.all_gather_object(gathered, (keys, counts))
        if rank == 0:
            allk = np.concatenate([g[0] for g in gathered])  # one round: rank order = code order
            allc = np.concatenate([g[1] for g in gathered])
            if int(os.environ.get("KC_SPARSE_RADIX_RBITS", "0")) > 0:  # several rounds: one ascending range per round and rank
                for g in gathered:
                    assert (np.diff(g[0].astype(np.int64)) > 0).all(), "a rank's keys must be ascending"
                order = np.argsort(allk, kind="stable")
                allk, allc = allk[order], allc[order]
            whole = O.gen_reads(0xB2000004, 20_000_000, 150, 200, 0, nreads)
            wk, wc, _ = O.count_sparse(whole, k)
            assert allk.size == wk.size and (allk == wk).all() and (allc == wc).all(), "sharded radix != oracle (k=%d)" % k
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("NCCL_RADIX_WORKER_OK world=%d" % world)
        sys.stdout.flush()
    os._exit(0)  # no library teardown (see bench.py: leave())


if __name__ == "__main__":
    main()

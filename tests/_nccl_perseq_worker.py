"""world_size-N GPU worker (torch.distributed.run, NCCL) for the per-sequence mode of SURVEY §8e: the sequences are
sharded across the ranks (cut at sequence starts, equal bytes), every rank counts its own with kc_count_per_seq, the
columns are all-gathered; every rank must hold the oracle's whole table int32[4^k][num_seqs], and the distance step
run on it must give the oracle's distances.  tests/test_sharding_gloo.py runs this file unmodified on the CPU
(emulator library + torch stand-ins + gloo)."""
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
    # KC_NCCL_PERSEQ_CASES="nseqs:maxlen:k,..." (smaller cases for the CPU dry run)
    cases = [tuple(int(x) for x in c.split(":")) for c in os.environ.get("KC_NCCL_PERSEQ_CASES", "400:200000:3,97:50000:6,5:3000:4").split(",")]
    for nseqs, maxlen, k in cases:
        rng = np.random.default_rng(100 + nseqs)   # the same sequences on every rank; a rank uploads only its own
        alphabet = np.frombuffer(b"ACGTACGTACGTACGTN", dtype=np.uint8)
        seqs = [alphabet[rng.integers(0, alphabet.size, int(rng.integers(1, maxlen)))] for _ in range(nseqs)]
        data = np.concatenate([np.concatenate([s, np.zeros(1, np.uint8)]) for s in seqs])   # main.cu:537-543
        offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)

        def count_seqs(s0, s1):
            if s1 == s0:
                return torch.zeros((kmerb200.num_kmers(k), 0), dtype=torch.int32, device=dev)
            mine = np.ascontiguousarray(data[offs[s0]:offs[s1]])
            d = torch.zeros(mine.size + 64, dtype=torch.uint8, device=dev)
            d[: mine.size] = torch.from_numpy(mine)
            d_off = torch.from_numpy(offs[s0:s1 + 1] - offs[s0]).to(dev)
            return ctx.count_per_seq(d, d_off, s1 - s0, k)

        sums = D.count_per_seq_sharded(count_seqs, offs, rank, world)
        torch.cuda.synchronize()
        want, _ = O.count_per_seq(data, offs, k)
        assert sums.shape == (kmerb200.num_kmers(k), nseqs), sums.shape
        assert (sums.cpu().numpy() == want).all(), "per-seq sharded != oracle (rank %d, %d seqs, k=%d)" % (rank, nseqs, k)
        # the distance step on the gathered table (any rank can run it: all hold all columns)
        d_offs = torch.from_numpy(offs).to(dev)
        got = ctx.kmer_distance(sums, d_offs, nseqs, k).cpu().numpy()
        assert got.tobytes() == O.distance(want, offs, k).tobytes(), "distances differ (rank %d)" % rank
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("NCCL_PERSEQ_WORKER_OK world=%d" % world)
        sys.stdout.flush()
    os._exit(0)  # no library teardown (see bench.py: leave())


if __name__ == "__main__":
    main()

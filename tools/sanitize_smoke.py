"""Small run of every kernel family, meant for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`
(closed on the round-2 pool: there the script only runs plain; tests/emu under AddressSanitizer covers the bounds on the CPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200")):
    sys.path.insert(0, p)
import kmerb200  # noqa: E402
import oracle as O  # noqa: E402

ctx = kmerb200.Context(0)
n = 700_001
host = O.gen_genome(7, n, 3, 300, 12, 0, n)
buf = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
buf[3:3 + n] = torch.from_numpy(host)
ptr = buf.data_ptr() + 3  # unaligned on purpose
# the radix path (scatter, leaf pass with its shared-memory hash table, scan, gather); KC_SPARSE_RADIX_RBITS=2 in the
# environment adds the round filter and the appended rounds.  SANITIZE_ONLY=radix runs just this part.
reads = O.gen_reads(11, 40_000, 100, 50, 0, 3000)
d_reads = torch.from_numpy(reads).cuda()
for k in (13, 21, 31):
    keys, counts = ctx.count_sparse(d_reads, reads.size, k, kmerb200.SPARSE_RADIX | kmerb200.SPARSE_NO_FALLBACK).to_host()
    wk, wc, _ = O.count_sparse(reads, k)
    assert keys.size == wk.size and (keys == wk).all() and (counts == wc).all(), k
if os.environ.get("SANITIZE_ONLY") == "radix":
    print("sanitize_smoke (radix only) ok, launches", ctx.launch_count)
    ctx.close()
    sys.exit(0)
for k, algo in ((3, 0), (7, 0), (8, 0), (10, 2), (12, 2), (12, 1), (13, 0)):
    t = torch.zeros(4 ** k, dtype=torch.int32, device="cuda")
    ctx.count_dense_range(ptr, n, 0, n, k, t, algo=algo)
    torch.cuda.synchronize()
    want, _ = O.count_dense(host, k)
    assert (t.cpu().numpy().view(np.uint32) == want).all(), (k, algo)
big = ctx.gen_bases(5, 0, 4_600_000)   # k = 8 shared-memory path needs >= 2^22 windows
t = torch.zeros(4 ** 8, dtype=torch.int32, device="cuda")
ctx.count_dense_range(big, 4_600_000, 0, 4_600_000, 8, t)
torch.cuda.synchronize()
assert int(t.to(torch.int64).sum()) == 4_600_000 - 7
seqs = [host[:5000], host[5000:5003], host[6000:90000]]
data = np.concatenate([np.concatenate([s, np.zeros(1, np.uint8)]) for s in seqs])
offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)
d_offs = torch.from_numpy(offs).cuda()
sums = ctx.count_per_seq(torch.from_numpy(data).cuda(), d_offs, 3, 4)
want, _ = O.count_per_seq(data, offs, 4)
assert (sums.cpu().numpy() == want).all()
dist = ctx.kmer_distance(sums, d_offs, 3, 4).cpu().numpy()
assert np.array_equal(dist, O.distance(want, offs, 4), equal_nan=True)  # a sequence shorter than k gives 0/0 on both sides
for algo in (0, 1):
    for k in (17, 31):
        keys, counts = ctx.count_sparse(ptr, 200_000, k, algo).to_host()
        wk, wc, _ = O.count_sparse(host[:200_000], k)
        assert (keys == wk).all() and (counts == wc).all()
h = ctx.count_dense_host(host, 12)
assert (h == O.count_dense(host, 12)[0]).all()
print("sanitize_smoke ok, launches", ctx.launch_count)
ctx.close()

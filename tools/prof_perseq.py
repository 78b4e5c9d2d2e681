"""one perseq_kernel launch per k for ncu (tools/r02_call31.sh): 10 000 sequences x 30 kb"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dna-kmeres-parallel_b200"))
import torch
import kmerb200 as K
ctx = K.Context(0)
nseq, slen = 10_000, 30_000
total = nseq * (slen + 1)
data = ctx.gen_bases(0xA11, 0, total)
ctx.synchronize()
offs = torch.arange(0, nseq + 1, dtype=torch.int64, device="cuda:0") * (slen + 1)
data[offs[1:] - 1] = 0
for k in (3, 6):
    sums = torch.zeros((K.num_kmers(k), nseq), dtype=torch.int32, device="cuda:0")
    for _ in range(3):
        ctx.count_per_seq(data, offs, nseq, k, sums=sums)
    print(k, int(sums.sum().item()))

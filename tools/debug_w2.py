"""GPU debug aid: KC_DENSE_PARTITION_WIDE2 vs KC_DENSE_PARTITION (sum of counts, differing bins)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-kmeres-parallel_b200")):
    sys.path.insert(0, p)
import torch
import kmerb200 as K

ctx = K.Context(0)
L = 1 << 24
data = ctx.gen_genome(0xB2000003, L, 0, 0, 12, 0, L)
a = torch.zeros(K.num_kmers(12), dtype=torch.int32, device="cuda:0")
ctx.count_dense_range(data, L, 0, L, 12, a, algo=K.DENSE_PARTITION)
for rep in range(4):
    b = torch.zeros_like(a)
    ctx.count_dense_range(data, L, 0, L, 12, b, algo=K.DENSE_PARTITION_WIDE2)
    torch.cuda.synchronize()
    d = (b.to(torch.int64) - a.to(torch.int64))
    nz = torch.nonzero(d).flatten()
    print("rep %d: differing bins %d" % (rep, nz.numel()))
    plus = sorted((int(i), int(d[i])) for i in nz.tolist() if d[i] > 0)
    minus = sorted((int(i), int(d[i])) for i in nz.tolist() if d[i] < 0)
    print("  plus ", [(hex(i), v) for i, v in plus][:42])
    print("  minus", [(hex(i), v) for i, v in minus][:42])
ctx.close()

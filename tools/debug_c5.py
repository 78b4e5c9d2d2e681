"""GPU debug aid: the flow of tests/test_gpu_parity.py::test_config5_quarter_scale_self_check with KC_TRACE"""
import os, sys
os.environ["KC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-kmeres-parallel_b200")):
    sys.path.insert(0, p)
import torch
import kmerb200 as K
ctx = K.Context(0)
nreads, k = 50_000_000, 31
if len(sys.argv) > 1:
    ctx.release_memory()
    torch.cuda.empty_cache()
reads = ctx.gen_reads(0xB2000005, 1_000_000_000, 150, 200, 0, nreads)
nb = nreads * 151
print("free", torch.cuda.mem_get_info())
fin = ctx.window_fingerprint(reads, nb, k)
print(fin)
try:
    sp = ctx.count_sparse(reads, nb, k, K.SPARSE_RADIX | K.SPARSE_NO_FALLBACK)
    print(len(sp), ctx.sparse_fingerprint(sp))
except Exception as e:
    print("ERR", e)

"""How fast can 8 B200s merge a uint32[4^12] table?  reduce vs all_reduce vs reduce_scatter (NCCL)."""
import os
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << 24
t = torch.ones(n, dtype=torch.int32, device="cuda")
out = torch.empty(n // world, dtype=torch.int32, device="cuda")


def timeit(fn, name):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-16s world=%d  %.3f ms" % (name, world, ms.item()), flush=True)


timeit(lambda: dist.reduce(t, dst=0), "reduce")
timeit(lambda: dist.all_reduce(t), "all_reduce")
timeit(lambda: dist.reduce_scatter_tensor(out, t), "reduce_scatter")
tf = t.view(torch.float32)
timeit(lambda: dist.all_reduce(tf), "all_reduce f32")
dist.destroy_process_group()

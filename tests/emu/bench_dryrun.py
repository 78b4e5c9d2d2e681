"""bench.py's GPU arm on a box without a GPU — TEST INFRASTRUCTURE ONLY.

The Python of bench.py's GPU arm (variant probes in child processes, timed region, roofline, e2e through
both host paths, CPU baseline, the JSON line) can otherwise only run at round end on a B200, where a typo
costs the round's measurement.  This runner executes that very code with
  * libkmerb200_emu.so (the product's kernels compiled against the SIMT emulator, tests/emu/) in the place
    of libkmerb200.so, and
  * torch's CUDA surface replaced by CPU stand-ins (device arguments dropped, streams / events / synchronize
    as no-ops, graph capture refusing so that bench.py takes its plain-launch path),
on a tiny workload.  Nothing here is measured or shipped; the numbers of such a run mean nothing and the line
says so (`"data": "... DRY RUN"` is added by the caller's checks, not by bench.py).

    python tests/emu/bench_dryrun.py --workload tiny_k12 --steps 1 [...bench.py flags]
"""
import contextlib
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def install_shim():
    import torch

    def strip(kw):
        kw.pop("pin_memory", None)
        if "device" in kw:
            kw["device"] = "cpu"
        return kw

    for name in ("zeros", "empty", "full", "arange", "tensor", "ones"):
        orig = getattr(torch, name)

        def make(orig):
            return lambda *a, **kw: orig(*a, **strip(kw))
        setattr(torch, name, make(orig))

    orig_to = torch.Tensor.to

    def to(self, *a, **kw):
        a = tuple("cpu" if (isinstance(x, torch.device) and x.type == "cuda") or (isinstance(x, str) and x.startswith("cuda")) else x
                  for x in a)
        if "device" in kw:
            kw["device"] = "cpu"
        return orig_to(self, *a, **kw)
    torch.Tensor.to = to
    torch.Tensor.pin_memory = lambda self, *a, **kw: self
    torch.Tensor.cuda = lambda self, *a, **kw: self

    class Stream:
        cuda_stream = 0

        def synchronize(self):
            pass

        def wait_stream(self, other):
            pass

    class Event:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

        def synchronize(self):
            pass

    class NoGraph:
        def __init__(self, *a, **kw):
            raise RuntimeError("no CUDA graphs in the dry run")

    cur = Stream()
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **kw: None
    torch.cuda.current_stream = lambda *a, **kw: cur
    torch.cuda.device_count = lambda: 1
    torch.cuda.Stream = Stream
    torch.cuda.Event = Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.CUDAGraph = NoGraph

    # N > 1: the same collectives over gloo on CPU tensors
    import torch.distributed as dist
    orig_init = dist.init_process_group

    def init_pg(backend=None, **kw):
        kw.pop("device_id", None)
        return orig_init("gloo", **kw)
    dist.init_process_group = init_pg

    import emu_harness
    emu_harness.build()
    import kmerb200
    kmerb200.LIB_PATH = emu_harness.LIB_PATH
    ctx_init = kmerb200.Context.__init__
    kmerb200.Context.__init__ = lambda self, device=0: ctx_init(self, 0)  # the emulator has one device; every rank uses it


def main():
    install_shim()
    import bench
    bench.PROBE_SCRIPT = os.path.abspath(__file__)  # the probes' child processes need the same stand-ins
    return bench.main()


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Times the rows of SURVEY §8 that bench.py does not carry (one B200, CUDA events, inputs resident unless noted):
the reference-shaped per-sequence table (kc_count_per_seq, kernels.h:113-144 layout), the distance step
(kernels.h:85-109), FASTA ingest (host threads vs the GPU transducer) and the 2-bit packed store.
Prints one JSON object; profiles/r02_aux_rows.json is its output on the round's box.
usage: python tools/measure_aux.py > gpurun_out/r02_aux_rows.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dna-kmeres-parallel_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import kmerb200 as K  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ctx = K.Context(0)
    dev = "cuda:0"
    out = {"device": torch.cuda.get_device_name(0), "rows": []}

    # ---- per-sequence table + distances: n sequences of ~len bases (the reference's plasmid set is 1000-odd x ~30 kb)
    for nseq, slen in ((1000, 30_000), (10_000, 30_000)):
        total = nseq * (slen + 1)
        data = ctx.gen_bases(0xA11, 0, total)
        ctx.synchronize()
        offs = torch.arange(0, nseq + 1, dtype=torch.int64, device=dev) * (slen + 1)
        data[offs[1:] - 1] = 0  # terminators
        for k in (3, 6, 8):
            sums = torch.zeros((K.num_kmers(k), nseq), dtype=torch.int32, device=dev)

            def run():
                sums.zero_()
                ctx.count_per_seq(data, offs, nseq, k, sums=sums, sync=False)
            ms = timed(run)
            row = {"row": "a1 per-sequence table", "k": k, "num_seqs": nseq, "seq_len": slen, "ms": ms,
                   "gbases_per_s": nseq * slen / ms / 1e6,
                   "algorithmic_GBps": (total + 4.0 * K.num_kmers(k) * nseq) / ms / 1e6}
            out["rows"].append(row)
            if nseq <= 1000 and k <= 6:
                ctx.count_per_seq(data, offs, nseq, k, sums=sums)
                dist = torch.zeros(nseq * (nseq - 1) // 2, dtype=torch.float32, device=dev)
                msd = timed(lambda: ctx.kmer_distance(sums, offs, nseq, k, dist=dist))
                out["rows"].append({"row": "f1 distance", "k": k, "num_seqs": nseq, "ms": msd,
                                    "mpairs_per_s": nseq * (nseq - 1) / 2 / msd / 1e3})
        del data, offs

    # ---- FASTA ingest: a file image of 20 000 records x 50 kb, 70 columns per line
    rng = np.random.default_rng(7)
    nrec, rlen, width = 20_000, 50_000, 70
    line = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, rlen)]
    body = np.insert(line, np.arange(width, rlen, width), ord("\n"))
    rec = np.concatenate([np.frombuffer(b">seq|id with words\n", dtype=np.uint8), body, np.frombuffer(b"\n\n", dtype=np.uint8)])
    image = np.tile(rec, nrec)
    nb = image.size
    raw = image.tobytes()
    for threads in (1, 0):
        ms = wall(lambda: K.SeqSet.from_memory_threads(raw, K.IMPORT_BLANKLINE, threads).close())
        out["rows"].append({"row": "a5/f2 FASTA ingest, host threads", "threads": threads or os.cpu_count(), "bytes": nb, "ms": ms,
                            "GBps": nb / ms / 1e6})
    d_raw = torch.from_numpy(image).to(dev)
    torch.cuda.synchronize()
    ms = wall(lambda: K.SeqSet.from_device(ctx, d_raw, raw, nb, K.IMPORT_BLANKLINE).close())
    out["rows"].append({"row": "f2 FASTA ingest, GPU transducer (image resident, host copy of the result included)", "bytes": nb, "ms": ms,
                        "GBps": nb / ms / 1e6})
    del d_raw

    # ---- 2-bit packed store: pack on the GPU, count from the packed form
    L, k = 1_000_000_000, 12
    data = ctx.gen_genome(0xB2000003, L, 300, 3000, k, 0, L)
    ctx.synchronize()
    packed, mask = ctx.pack_2bit(data, L)
    ms = timed(lambda: ctx.pack_2bit(data, L))
    out["rows"].append({"row": "f4 pack to 2 bits + validity bitmap (GPU)", "bases": L, "ms": ms, "gbases_per_s": L / ms / 1e6})
    table = torch.zeros(K.num_kmers(k), dtype=torch.int32, device=dev)
    ms = timed(lambda: ctx.count_dense_packed(packed, mask, L, k, table=table))
    out["rows"].append({"row": "f4 count from the packed store (0.375 B/base at rest)", "bases": L, "k": k, "ms": ms,
                        "gbases_per_s": L / ms / 1e6})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""ctypes access to libkmerb200_emu.so — the product's kernels compiled for the CPU against the
test-only SIMT emulator (tests/emu/simt_emu.h).  TEST INFRASTRUCTURE ONLY: it checks kernel
LOGIC against the oracle where there is no GPU; it is never the thing measured or shipped.

The emulator reads KC_EMU_SEED (0 = deterministic round-robin, != 0 = random fiber
scheduling with preemption inside shared-memory/atomic helpers) and KC_EMU_SMS once per
process, so tests that want several seeds run this module as a subprocess:

    python tests/emu/emu_harness.py <case> [args...]      (see main() below)
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "_build", "libkmerb200_emu.so")
_LIB = None

KC_DENSE_AUTO, KC_DENSE_DIRECT, KC_DENSE_PARTITION = 0, 1, 2
KC_SPARSE_HASH, KC_SPARSE_SORT, KC_SPARSE_RADIX = 0, 1, 2
KC_SPARSE_UNSORTED = 0x100
KC_SPARSE_NO_FALLBACK = 0x200


def build():
    subprocess.run(["make", "-s", "-j8", "-C", HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.kc_last_error.restype = C.c_char_p
        L.kc_last_error.argtypes = [C.c_void_p]
        L.kc_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.kc_ctx_destroy.argtypes = [C.c_void_p]
        L.kc_device_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.kc_device_free.argtypes = [C.c_void_p, C.c_void_p]
        L.kc_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.kc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.kc_memset_d.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
        L.kc_count_dense_range_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                                 C.c_void_p, C.c_int, C.c_void_p]
        L.kc_count_per_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.kc_count_sparse.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64,
                                      C.POINTER(C.c_void_p)]
        L.kc_sparse_free.argtypes = [C.c_void_p]
        L.kc_sparse_size.restype = C.c_uint64
        L.kc_sparse_size.argtypes = [C.c_void_p]
        L.kc_sparse_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kc_kmer_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.kc_gen_bases.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.kc_count_dense_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.kc_count_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.kc_import_seqs_device.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
        L.kc_seqset_to_device.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.kc_import_seqs_gpu.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.kc_import_seqs_mem.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_long, C.POINTER(C.c_void_p)]
        L.kc_seqset_free.argtypes = [C.c_void_p]
        L.kc_seqset_num_seqs.restype = C.c_uint32
        L.kc_seqset_num_seqs.argtypes = [C.c_void_p]
        L.kc_seqset_num_ids.restype = C.c_uint32
        L.kc_seqset_num_ids.argtypes = [C.c_void_p]
        L.kc_seqset_nbytes.restype = C.c_uint64
        L.kc_seqset_nbytes.argtypes = [C.c_void_p]
        L.kc_seqset_data.restype = C.c_void_p
        L.kc_seqset_data.argtypes = [C.c_void_p]
        L.kc_seqset_offsets.restype = C.c_void_p
        L.kc_seqset_offsets.argtypes = [C.c_void_p]
        L.kc_seqset_id.restype = C.c_char_p
        L.kc_seqset_id.argtypes = [C.c_void_p, C.c_uint32]
        L.kc_packed_bytes.restype = C.c_uint64
        L.kc_packed_bytes.argtypes = [C.c_uint64]
        L.kc_badmask_bytes.restype = C.c_uint64
        L.kc_badmask_bytes.argtypes = [C.c_uint64]
        L.kc_pack_2bit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kc_unpack_2bit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.kc_count_dense_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.kc_ctx_last_h2d_bytes.restype = C.c_uint64
        L.kc_ctx_last_h2d_bytes.argtypes = [C.c_void_p]
        L.kc_pack_2bit_host_body.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.kc_count_dense_host_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.kc_count_dense_host_packed_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int]
        L.kc_sparse_radix_plan.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p]
        L.kc_sparse_radix_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kc_sparse_radix_scatter_round.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.kc_sparse_radix_count_round.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                  C.c_void_p]
        L.kc_sparse_radix_count_round_append.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                  C.c_void_p]
        L.kc_sparse_radix_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.POINTER(C.c_void_p)]
        L.kc_gen_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64,
                                    C.c_uint64, C.c_void_p, C.c_void_p]
        L.kc_gen_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                   C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class RadixPlan(C.Structure):  # kc_radix_plan (include/kmer_b200.h)
    _fields_ = [("k", C.c_int32), ("world", C.c_uint32), ("partitions", C.c_uint32), ("parts_per_rank", C.c_uint32),
                ("grid", C.c_uint32), ("rec_bytes", C.c_uint32), ("shape", C.c_uint32), ("round_bits", C.c_uint32),
                ("max_windows", C.c_uint64), ("region_records", C.c_uint64), ("slab_bytes", C.c_uint64),
                ("counts_bytes", C.c_uint64)]


class EmuContext:
    def __init__(self):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.kc_ctx_create(0, C.byref(h))
        assert rc == 0, self.L.kc_last_error(None)
        self.h = h

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("emu rc=%d: %s" % (rc, self.L.kc_last_error(self.h).decode()))

    def alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.L.kc_device_alloc(self.h, nbytes, C.byref(p)))
        return p

    def free(self, p):
        self.L.kc_device_free(self.h, p)

    def upload(self, arr, offset=0):
        """device copy of a numpy array, placed `offset` bytes into its allocation"""
        a = np.ascontiguousarray(arr)
        base = self.alloc(a.nbytes + offset)
        p = C.c_void_p(base.value + offset)
        if a.nbytes:
            self.check(self.L.kc_memcpy_h2d(self.h, p, a.ctypes.data, a.nbytes))
        return base, p

    def download(self, p, nbytes, dtype):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if nbytes:
            self.check(self.L.kc_memcpy_d2h(self.h, out.ctypes.data, p, nbytes))
        return out

    def count_dense_range(self, data, k, wb=None, we=None, algo=KC_DENSE_AUTO, offset=0):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        n = a.size
        nwin = max(0, n - k + 1)
        wb = 0 if wb is None else wb
        we = nwin if we is None else we
        base, p = self.upload(a, offset)
        nb = 4 << (2 * k)
        t = self.alloc(nb)
        self.check(self.L.kc_memset_d(self.h, t, 0, nb))
        self.check(self.L.kc_count_dense_range_async(self.h, p, n, wb, we, k, t, algo, None))
        out = self.download(t, nb, np.uint32)
        self.free(t)
        self.free(base)
        return out

    def count_sparse(self, data, k, algo, hint=0, offset=0):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        base, p = self.upload(a, offset)
        sp = C.c_void_p()
        self.check(self.L.kc_count_sparse(self.h, p, a.size, k, algo, hint, C.byref(sp)))
        n = int(self.L.kc_sparse_size(sp))
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        self.check(self.L.kc_sparse_copy_to_host(self.h, sp, keys.ctypes.data, counts.ctypes.data))
        self.L.kc_sparse_free(sp)
        self.free(base)
        return keys, counts

    def close(self):
        self.L.kc_ctx_destroy(self.h)


def _oracle():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    O.lib()
    return O


def make_input(kind, n, seed, k):
    """Test inputs: 'genome' = the bench generator with N runs; 'dirty' = random bytes from a
    small alphabet with separators, lower case and N; 'polyA' = one k-mer everywhere."""
    O = _oracle()
    if kind == "genome":
        return O.gen_genome(seed, n, 3, 40, k, 0, n)
    if kind == "polyA":
        return np.full(n, ord("A"), dtype=np.uint8)
    if kind.startswith("sparseN"):  # random ACGT with an invalid byte about every <period> bases
        period = int(kind[7:])
        rng = np.random.default_rng(seed)
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        pos = np.flatnonzero(rng.random(n) < 1.0 / period)
        a[pos] = np.frombuffer(b"N\n\0a", dtype=np.uint8)[rng.integers(0, 4, pos.size)]
        return a
    if kind == "runsT":  # random ACGT with runs of T: windows whose code is all ones (the leaf table's EMPTY value)
        rng = np.random.default_rng(seed)
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        for pos in rng.integers(0, max(n - 2 * k, 1), 5):  # (few and short: long runs are skew, the regions' business)
            a[pos:pos + k + int(rng.integers(0, 6))] = ord("T")
        return a
    if kind == "fewA":  # random ACGT with 21 % A: the round of the codes that END in A is smaller than the others
        rng = np.random.default_rng(seed)
        return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, n, p=[0.21, 0.2633, 0.2633, 0.2634])].copy()
    if kind == "skew":  # 90 % of the windows fall into a few partitions
        rng = np.random.default_rng(seed)
        a = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
        a = a.copy()
        mask = rng.random(n) < 0.9
        a[mask] = ord("C")
        return a
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTNacgt\n\0|>", dtype=np.uint8)
    return alpha[rng.integers(0, alpha.size, n)]


def case_dense(args):
    k, n, algo, kind, seed, offset = int(args[0]), int(args[1]), int(args[2]), args[3], int(args[4]), int(args[5])
    O = _oracle()
    data = make_input(kind, n, seed, k)
    ctx = EmuContext()
    got = ctx.count_dense_range(data, k, algo=algo, offset=offset)
    want, _ = O.count_dense(data, k)
    assert (got == want).all(), "dense k=%d algo=%d %s n=%d: %d bins differ" % (k, algo, kind, n, int((got != want).sum()))
    # a window sub-range too (the multi-GPU shard form)
    wb, we = n // 3, n - n // 5
    got = ctx.count_dense_range(data, k, wb, we, algo=algo, offset=offset)
    want = O.count_dense_range(data, k, wb, we)
    if isinstance(want, tuple):
        want = want[0]
    assert (got == want).all(), "dense range k=%d algo=%d: %d bins differ" % (k, algo, int((got != want).sum()))
    ctx.close()
    print("ok dense", *args)


def case_sparse(args):
    k, n, algo, kind, seed, offset = int(args[0]), int(args[1]), int(args[2]), args[3], int(args[4]), int(args[5])
    O = _oracle()
    if kind == "reads":      # deep coverage of a tiny genome: few distinct k-mers, many repeats
        nreads = n // 101
        data = O.gen_reads(seed, 5000, 100, 50, 0, nreads)
    elif kind == "readsU":   # shallow coverage: mostly distinct k-mers, uniform over the partitions
        nreads = n // 101
        data = O.gen_reads(seed, 4 * n, 100, 50, 0, nreads)
    else:
        data = make_input(kind, n, seed, k)
    ctx = EmuContext()
    keys, counts = ctx.count_sparse(data, k, algo, offset=offset)
    if algo & KC_SPARSE_UNSORTED:
        order = np.argsort(keys, kind="stable")
        keys, counts = keys[order], counts[order]
    wk, wc, _ = O.count_sparse(data, k)
    assert keys.size == wk.size, "sparse k=%d algo=%d: %d distinct, oracle %d" % (k, algo, keys.size, wk.size)
    assert (keys == wk).all() and (counts == wc).all(), "sparse k=%d algo=%d differs from the oracle" % (k, algo)
    ctx.close()
    print("ok sparse", *args, "distinct", keys.size)


def case_perseq(args):
    """reference-shaped per-sequence table sums[4^k][n] (kernels.h:113-144 layout) + the distance step"""
    k, nseq, seed = int(args[0]), int(args[1]), int(args[2])
    O = _oracle()
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTACGTACGTACGTNa", dtype=np.uint8)
    seqs = [alpha[rng.integers(0, alpha.size, int(rng.integers(0, 40_000 if i % 7 == 0 else 300)))] for i in range(nseq)]
    data = np.concatenate([np.concatenate([s, np.zeros(1, np.uint8)]) for s in seqs])
    offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)
    ctx = EmuContext()
    b1, d_data = ctx.upload(data, 5)
    b2, d_off = ctx.upload(offs)
    nb = (4 << (2 * k)) * nseq
    d_sums = ctx.alloc(nb)
    ctx.check(ctx.L.kc_count_per_seq(ctx.h, d_data, d_off, nseq, k, d_sums))
    got = ctx.download(d_sums, nb, np.int32).reshape(4 ** k, nseq)
    want, _ = O.count_per_seq(data, offs, k)
    assert (got == want).all(), "per-seq k=%d: %d cells differ" % (k, int((got != want).sum()))
    npairs = nseq * (nseq - 1) // 2
    d_dist = ctx.alloc(4 * max(npairs, 1))
    ctx.check(ctx.L.kc_kmer_distance(ctx.h, d_sums, d_off, nseq, k, d_dist))
    dist = ctx.download(d_dist, 4 * npairs, np.float32)
    wd = O.distance(want, offs, k)
    assert dist.tobytes() == np.asarray(wd, dtype=np.float32).tobytes(), "distance differs"
    ctx.close()
    print("ok perseq", *args)


def case_gen(args):
    """the HBM generators against the oracle's (same bytes on both sides)"""
    n, seed = int(args[0]), int(args[1])
    O = _oracle()
    ctx = EmuContext()
    d = ctx.alloc(n + 3)
    p = C.c_void_p(d.value + 3)  # misaligned destination: the scalar tail path
    ctx.check(ctx.L.kc_gen_genome(ctx.h, seed, 10 * n, 7, 90, 12, 3 * n, n, p, None))
    assert (ctx.download(p, n, np.uint8) == O.gen_genome(seed, 10 * n, 7, 90, 12, 3 * n, n)).all()
    ctx.check(ctx.L.kc_gen_bases(ctx.h, seed, 123, n, d, None))
    assert (ctx.download(d, n, np.uint8) == O.gen_bases(seed, 123, n)).all()
    nreads = n // 151
    d2 = ctx.alloc(nreads * 151)
    ctx.check(ctx.L.kc_gen_reads(ctx.h, seed, 50_000, 150, 200, 17, nreads, d2, None))
    assert (ctx.download(d2, nreads * 151, np.uint8) == O.gen_reads(seed, 50_000, 150, 200, 17, nreads)).all()
    ctx.close()
    print("ok gen", *args)


def case_radix_sharded(args):
    """The multi-GPU radix path with the ranks emulated one after the other: every rank scatters
    its reads, the slab blocks of each rank's partition range are exchanged (numpy slicing stands
    in for the all-to-all), every rank counts what it owns; the concatenation in rank order must
    be the oracle's sorted result for ALL the reads."""
    k, nreads, world, seed = int(args[0]), int(args[1]), int(args[2]), int(args[3])
    O = _oracle()
    ctx = EmuContext()
    L = ctx.L
    shards = [(nreads * r // world, nreads * (r + 1) // world) for r in range(world)]
    reads = [O.gen_reads(seed, 8 * nreads, 100, 50, r0, r1 - r0) for r0, r1 in shards]
    plan = RadixPlan()
    ctx.check(L.kc_sparse_radix_plan(ctx.h, max(max(x.size - k + 1, 0) for x in reads), k, world, C.byref(plan)))
    assert plan.parts_per_rank * world == plan.partitions
    rounds = 1 << plan.round_bits  # > 1 only when KC_SPARSE_RADIX_RBITS forces it at these sizes
    sb, cb = plan.slab_bytes // world, plan.counts_bytes // 4 // world   # one rank's block of a scatter output
    keys, cnts = [[] for _ in range(world)], [[] for _ in range(world)]
    for rnd in range(rounds):
        slabs, counts = [], []
        for x in reads:
            base, p = ctx.upload(x, 3)
            d_s, d_c = ctx.alloc(plan.slab_bytes), ctx.alloc(plan.counts_bytes)
            ctx.check(L.kc_sparse_radix_scatter_round(ctx.h, p, x.size, C.byref(plan), rnd, d_s, d_c))
            slabs.append(ctx.download(d_s, plan.slab_bytes, np.uint8))
            counts.append(ctx.download(d_c, plan.counts_bytes, np.uint32))
            for q in (base, d_s, d_c):
                ctx.free(q)
        for o in range(world):
            recv_s = np.concatenate([slabs[src][o * sb:(o + 1) * sb] for src in range(world)])
            recv_c = np.concatenate([counts[src][o * cb:(o + 1) * cb] for src in range(world)])
            b1, d_s = ctx.upload(recv_s)
            b2, d_c = ctx.upload(recv_c)
            sp = C.c_void_p()
            ctx.check(L.kc_sparse_radix_count_round(ctx.h, C.byref(plan), rnd, d_s, d_c, world, o * plan.parts_per_rank, plan.parts_per_rank,
                                                    C.byref(sp)))
            n = int(L.kc_sparse_size(sp))
            kk, cc = np.empty(n, np.uint64), np.empty(n, np.uint32)
            ctx.check(L.kc_sparse_copy_to_host(ctx.h, sp, kk.ctypes.data, cc.ctypes.data))
            L.kc_sparse_free(sp)
            ctx.free(b1)
            ctx.free(b2)
            keys[o].append(kk)
            cnts[o].append(cc)
    # round-major, rank-minor is code order (one round: rank order)
    allk = np.concatenate([keys[o][rnd] for rnd in range(rounds) for o in range(world)])
    allc = np.concatenate([cnts[o][rnd] for rnd in range(rounds) for o in range(world)])
    keys = [np.concatenate(kl) for kl in keys]
    wk, wc, _ = O.count_sparse(np.concatenate(reads), k)
    assert allk.size == wk.size and (allk == wk).all() and (allc == wc).all(), "sharded radix differs from the oracle"
    assert all(kk.size > 0 for kk in keys), "a rank owns nothing?"
    ctx.close()
    print("ok radix_sharded", *args, "distinct per rank", [int(kk.size) for kk in keys])


def case_dense_host(args):
    """kc_count_dense_host: host buffer in, host table out (chunked staging + counting of the windows
    that END inside each chunk), and kc_count_dense (zeroing entry point)"""
    k, n, seed = int(args[0]), int(args[1]), int(args[2])
    O = _oracle()
    data = make_input("genome", n, seed, k)
    ctx = EmuContext()
    table = np.full(4 ** k, 0xDEADBEEF, dtype=np.uint32)
    ctx.check(ctx.L.kc_count_dense_host(ctx.h, data.ctypes.data, n, k, table.ctypes.data))
    want, _ = O.count_dense(data, k)
    assert (table == want).all(), "kc_count_dense_host differs"
    base, p = ctx.upload(data, 9)
    t = ctx.alloc(4 << (2 * k))
    ctx.check(ctx.L.kc_memset_d(ctx.h, t, 0x5A, 4 << (2 * k)))  # kc_count_dense must overwrite, not add
    ctx.check(ctx.L.kc_count_dense(ctx.h, p, n, k, t))
    assert (ctx.download(t, 4 << (2 * k), np.uint32) == want).all(), "kc_count_dense differs"
    ctx.close()
    print("ok dense_host", *args)


def case_dense_host_packed(args):
    """kc_count_dense_host_packed: host threads pack into the pinned ring, slots are copied, unpacked and
    counted behind the copies; KC_HOSTPACK_ITEM (env) makes the items small enough for ring wrap-arounds"""
    k, n, seed, kind, nthreads = int(args[0]), int(args[1]), int(args[2]), args[3], int(args[4])
    O = _oracle()
    data = make_input(kind, n, seed, k)
    ctx = EmuContext()
    want, _ = O.count_dense(data, k)
    for rep in range(2):  # the second call reuses the ring and the device image
        table = np.full(4 ** k, 0xDEADBEEF, dtype=np.uint32)
        ctx.check(ctx.L.kc_count_dense_host_packed(ctx.h, data.ctypes.data if n else None, n, k, table.ctypes.data, nthreads))
        assert (table == want).all(), "kc_count_dense_host_packed differs"
        sent = int(ctx.L.kc_ctx_last_h2d_bytes(ctx.h))
        if n >= k:  # per slot: packed bytes + 256-byte header + its dirty bitmap blocks (sparse) or its whole bitmap (full)
            item = int(os.environ.get("KC_HOSTPACK_ITEM", "0")) or (1 << 20)
            item = (item + 31) // 32 * 32 if item <= 1024 else (item + 1023) // 1024 * 1024
            item_words = item // 32
            block_words = 1 if item_words <= 32 else item_words // 32
            bpi = item_words // block_words
            slot, block = 16 * item, 32 * block_words
            valid = np.isin(data, np.frombuffer(b"ACGT", dtype=np.uint8))
            expect = 0
            for b in range(0, n, slot):
                e = min(b + slot, n)
                dirty = sum(1 for o in range(b, e, block) if not valid[o:min(o + block, e)].all() or (min(o + block, e) - o) % 32)
                expect += (e - b + 3) // 4 + 256
                expect += dirty * block_words * 4 if dirty <= (16 * bpi) // 4 else (e - b + 31) // 32 * 4
            assert sent == expect, (sent, expect)
    t = ctx.alloc(4 << (2 * k))   # the device-table variant overwrites whatever the table held
    ctx.check(ctx.L.kc_memset_d(ctx.h, t, 0x5A, 4 << (2 * k)))
    ctx.check(ctx.L.kc_count_dense_host_packed_dev(ctx.h, data.ctypes.data if n else None, n, k, t, nthreads))
    assert (ctx.download(t, 4 << (2 * k), np.uint32) == want).all(), "kc_count_dense_host_packed_dev differs"
    ctx.close()
    print("ok dense_host_packed", *args)


def case_packed(args):
    """f4: pack -> the reference sketch's layout (first base in the top two bits of each byte) + validity
    bitmap; unpack is the inverse up to invalid -> 'N'; counting from the store == counting the bytes"""
    k, n, seed = int(args[0]), int(args[1]), int(args[2])
    O = _oracle()
    data = make_input("dirty", n, seed, k)
    m = min(5000, n - n // 3)  # a clean stretch, so that large k sees valid windows too
    data[n // 3: n // 3 + m] = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(seed).integers(0, 4, m)]
    ctx = EmuContext()
    L = ctx.L
    base, p = ctx.upload(data)
    pb, mb = int(L.kc_packed_bytes(n)), int(L.kc_badmask_bytes(n))
    assert pb == (n + 3) // 4 and mb == (n + 31) // 32 * 4
    d_p, d_m = ctx.alloc(pb), ctx.alloc(mb)
    ctx.check(L.kc_pack_2bit(ctx.h, p, n, d_p, d_m, None))
    packed = ctx.download(d_p, pb, np.uint8)
    mask = ctx.download(d_m, mb, np.uint32)
    code = np.full(256, -1, dtype=np.int64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    c = code[data]
    bad = c < 0
    c2 = np.where(bad, 0, c).astype(np.uint8)
    pad = np.zeros((-n) % 4, dtype=np.uint8)
    q = np.concatenate([c2, pad]).reshape(-1, 4)
    want_packed = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]    # "AACG -> 00000110" (main.cu:83)
    assert (packed == want_packed).all(), "packed bytes differ"
    bits = np.concatenate([bad, np.ones((-n) % 32, dtype=bool)]).reshape(-1, 32)   # bases past the end count as invalid
    want_mask = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    assert (mask == want_mask).all(), "validity bitmap differs"
    # the host-side packer (hostpack.cpp: AVX-512 / AVX2 / scalar bodies, several threads) writes the same store
    for nthreads, body in ((1, 0), (3, 0), (1, 1), (2, 2)):
        hp = np.full(pb, 0xEE, dtype=np.uint8)
        hm = np.full(mb // 4, 0xEEEEEEEE, dtype=np.uint32)
        ctx.check(L.kc_pack_2bit_host_body(data.ctypes.data, n, hp.ctypes.data, hm.ctypes.data, nthreads, body))
        assert (hp == packed).all() and (hm == mask).all(), "host packer differs from pack_kernel (threads %d, body %d)" % (nthreads, body)
    d_o = ctx.alloc(n)
    ctx.check(L.kc_unpack_2bit(ctx.h, d_p, d_m, n, d_o, None))
    back = ctx.download(d_o, n, np.uint8)
    assert (back == np.where(bad, ord("N"), data)).all(), "unpack is not the inverse"
    t = ctx.alloc(4 << (2 * k))
    ctx.check(L.kc_count_dense_packed(ctx.h, d_p, d_m, n, k, t))
    want, _ = O.count_dense(data, k)
    assert (ctx.download(t, 4 << (2 * k), np.uint32) == want).all(), "count from the packed store differs"
    ctx.close()
    print("ok packed", *args)


def _seqset_tuple(L, h):
    n = int(L.kc_seqset_num_seqs(h))
    nb = int(L.kc_seqset_nbytes(h))
    dp = L.kc_seqset_data(h)
    data = C.string_at(dp, nb) if nb else b""
    offs = np.ctypeslib.as_array(C.cast(L.kc_seqset_offsets(h), C.POINTER(C.c_int64)), shape=(n + 1,)).tolist()
    ids = [L.kc_seqset_id(h, i) for i in range(int(L.kc_seqset_num_ids(h)))]
    return n, data, offs, ids


def case_ingest(args):
    """f2, device side: kc_import_seqs_device (seven-state byte transducer, per-tile maps composed by a
    scan) against the host loader kc_import_seqs_mem, on the reference-generated loader fixtures and on
    random files — with tiles of 16 bytes, so that lines, headers and records span many tiles"""
    import json
    seed, ntrials = int(args[0]), int(args[1])
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))
    texts = [c["fasta"].encode("latin-1") for c in golden["loader"]]
    rng = np.random.default_rng(seed)
    pieces = [b">h\n", b"ACGT\n", b"NNAC\n", b"\n", b"\r\n", b"acgt\n", b">x y\n", b"GG|TT\n", b"T", b"\n\n", b"CCC\r\n",
              b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n", b">\n", b"|\n", b"\r", b">"]
    for _ in range(ntrials):
        texts.append(b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=int(rng.integers(0, 80)))))
    # more tiles than the scan kernel has threads: every thread composes a RUN of tile maps
    texts.append(b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=9000)))
    ctx = EmuContext()
    L = ctx.L
    for text in texts:
        for mode in (0, 1):
            want = C.c_void_p()
            ctx.check(L.kc_import_seqs_mem(text, len(text), mode, 0, C.byref(want)))
            base, p = ctx.upload(np.frombuffer(text, dtype=np.uint8)) if text else (None, None)
            got = C.c_void_p()
            ctx.check(L.kc_import_seqs_device(ctx.h, p, text, len(text), mode, C.byref(got)))
            a, b = _seqset_tuple(L, want), _seqset_tuple(L, got)
            assert a == b, (text, mode, a, b)
            if a[0] and len(text) > 5000:  # the device-resident copy feeds the per-sequence count directly
                dd, do = C.c_void_p(), C.c_void_p()
                ctx.check(L.kc_seqset_to_device(ctx.h, got, C.byref(dd), C.byref(do)))
                d_sums = ctx.alloc(4 * 64 * a[0])
                ctx.check(L.kc_count_per_seq(ctx.h, dd, do, a[0], 3, d_sums))
                sums = ctx.download(d_sums, 4 * 64 * a[0], np.int32).reshape(64, a[0])
                ws, _ = _oracle().count_per_seq(np.frombuffer(a[1], dtype=np.uint8), np.array(a[2], dtype=np.int64), 3)
                assert (sums == ws).all(), "per-sequence counts of the GPU-parsed set differ"
                ctx.free(d_sums)
            L.kc_seqset_free(want)
            L.kc_seqset_free(got)
            if base is not None:
                ctx.free(base)
    # the file form, and its error text for a missing file
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".fasta") as f:
        f.write(texts[-1])
        f.flush()
        for mode in (0, 1):
            want, got = C.c_void_p(), C.c_void_p()
            ctx.check(L.kc_import_seqs_mem(texts[-1], len(texts[-1]), mode, 0, C.byref(want)))
            ctx.check(L.kc_import_seqs_gpu(ctx.h, f.name.encode(), mode, C.byref(got)))
            assert _seqset_tuple(L, want) == _seqset_tuple(L, got)
            L.kc_seqset_free(want)
            L.kc_seqset_free(got)
    got = C.c_void_p()
    assert L.kc_import_seqs_gpu(ctx.h, b"/nonexistent/all_seqs.fasta", 0, C.byref(got)) == -3
    assert b"Error opening" in L.kc_last_error(ctx.h)
    ctx.close()
    print("ok ingest", *args, "files", len(texts))


def case_fingerprint(args):
    """csrc/check.cu: the window fingerprint of an input == the fingerprint of its sparse and dense counts
    == the numpy value from the oracle's count (sum of count * mix64(code) mod 2^64), totals included"""
    k, n, kind, seed, offset = int(args[0]), int(args[1]), args[2], int(args[3]), int(args[4])
    O = _oracle()
    sys.path.insert(0, os.path.join(ROOT, "dna-kmeres-parallel_b200"))
    from kmerb200.distributed import mix64_np
    data = make_input(kind, n, seed, k)
    wk, wc, _ = O.count_sparse(data, k)
    with np.errstate(over="ignore"):
        want_fp = int((mix64_np(wk) * wc.astype(np.uint64)).sum(dtype=np.uint64))
    want_n = int(wc.astype(np.uint64).sum())
    ctx = EmuContext()
    L = ctx.L
    L.kc_window_fingerprint.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kc_sparse_fingerprint.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kc_dense_fingerprint.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    base, p = ctx.upload(np.ascontiguousarray(data, dtype=np.uint8), offset)
    fp, tot = C.c_uint64(), C.c_uint64()
    ctx.check(L.kc_window_fingerprint(ctx.h, p, data.size, k, C.byref(fp), C.byref(tot)))
    assert (fp.value, tot.value) == (want_fp, want_n), ("window", fp.value, tot.value, want_fp, want_n)
    sp = C.c_void_p()
    ctx.check(L.kc_count_sparse(ctx.h, p, data.size, k, 3, 0, C.byref(sp)))  # KC_SPARSE_AUTO
    desc = C.c_uint64(7)
    ctx.check(L.kc_sparse_fingerprint(ctx.h, sp, C.byref(fp), C.byref(tot), C.byref(desc)))
    assert (fp.value, tot.value, desc.value) == (want_fp, want_n, 0), ("sparse", fp.value, tot.value, desc.value, want_fp, want_n)
    L.kc_sparse_free(sp)
    if k <= 12:
        nb = 4 << (2 * k)
        t = ctx.alloc(nb)
        ctx.check(L.kc_memset_d(ctx.h, t, 0, nb))
        ctx.check(L.kc_count_dense_range_async(ctx.h, p, data.size, 0, max(0, data.size - k + 1), k, t, 0, None))
        ctx.check(L.kc_dense_fingerprint(ctx.h, t, k, C.byref(fp), C.byref(tot)))
        assert (fp.value, tot.value) == (want_fp, want_n), ("dense", fp.value, tot.value, want_fp, want_n)
        ctx.free(t)
    ctx.free(base)
    ctx.close()
    print("ok fingerprint", *args)


def case_radix_append(args):
    """kc_sparse_radix_count_round_append on one rank: the rounds of a plan (KC_SPARSE_RADIX_RBITS forces several) are
    appended to ONE result; with 'fewA' the first round is the smallest, so a later one does not fit the arrays that
    were sized from it and they grow (stat 14)."""
    k, n, kind, seed = int(args[0]), int(args[1]), args[2], int(args[3])
    O = _oracle()
    data = make_input(kind, n, seed, k) if kind != "readsU" else O.gen_reads(seed, 4 * n, 100, 50, 0, n // 101)
    ctx = EmuContext()
    L = ctx.L
    plan = RadixPlan()
    ctx.check(L.kc_sparse_radix_plan(ctx.h, max(data.size - k + 1, 0), k, 1, C.byref(plan)))
    base, p = ctx.upload(data, 3)
    d_s, d_c = ctx.alloc(plan.slab_bytes), ctx.alloc(plan.counts_bytes)
    acc = C.c_void_p()
    for rnd in range(1 << plan.round_bits):
        ctx.check(L.kc_sparse_radix_scatter_round(ctx.h, p, data.size, C.byref(plan), rnd, d_s, d_c))
        ctx.check(L.kc_sparse_radix_count_round_append(ctx.h, C.byref(plan), rnd, d_s, d_c, 1, 0, plan.partitions, C.byref(acc)))
    cnt = int(L.kc_sparse_size(acc))
    keys, counts = np.empty(cnt, np.uint64), np.empty(cnt, np.uint32)
    ctx.check(L.kc_sparse_copy_to_host(ctx.h, acc, keys.ctypes.data, counts.ctypes.data))
    L.kc_sparse_free(acc)
    for q in (base, d_s, d_c):
        ctx.free(q)
    wk, wc, _ = O.count_sparse(data, k)
    assert keys.size == wk.size and (keys == wk).all() and (counts == wc).all(), "appended rounds differ from the oracle (k=%d)" % k
    ctx.close()
    print("ok radix_append", *args, "rounds", 1 << plan.round_bits, "distinct", cnt)


CASES = {"radix_append": case_radix_append, "fingerprint": case_fingerprint, "dense_host_packed": case_dense_host_packed, "ingest": case_ingest, "packed": case_packed, "dense_host": case_dense_host, "radix_sharded": case_radix_sharded, "dense": case_dense, "sparse": case_sparse, "perseq": case_perseq, "gen": case_gen}

if __name__ == "__main__":
    CASES[sys.argv[1]](sys.argv[2:])

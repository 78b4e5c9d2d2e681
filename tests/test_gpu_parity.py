"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference goldens.  Every test here needs a B200: run with `-m gpu`."""
import os

import numpy as np
import pytest

from mt64 import mt19937_64

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def to_dev(arr):
    torch = _torch()
    a = np.frombuffer(arr, dtype=np.uint8) if isinstance(arr, (bytes, bytearray)) else np.ascontiguousarray(arr, dtype=np.uint8)
    # +64 slack so deliberately misaligned views stay inside the allocation
    buf = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda:0")
    if a.size:
        buf[: a.size] = torch.from_numpy(a.copy())
    return buf


def dense_gpu(ctx, kmerlib, dbuf, n, k, algo=0, ranges=None, offset=0):
    torch = _torch()
    table = torch.zeros(kmerlib.num_kmers(k), dtype=torch.int32, device="cuda:0")
    ptr = dbuf.data_ptr() + offset
    for b, e in (ranges or [(0, n)]):
        ctx.count_dense_range(ptr, n, b, e, k, table, algo=algo)
    torch.cuda.synchronize()
    return table.cpu().numpy().view(np.uint32)


def dirty(oracle, n, seed=1):
    s = oracle.gen_genome(seed, n, max(1, n // 50000), max(1, n // 500), 12, 0, n).copy()
    rng = np.random.default_rng(seed)
    junk = np.frombuffer(b"acgtn\r\n\0|-*RYKMX", dtype=np.uint8)
    idx = rng.integers(0, n, size=max(1, n // 300))
    s[idx] = junk[rng.integers(0, junk.size, size=idx.size)]
    return s


# --------------------------------------------------------------------------
def test_generators_match_oracle(ctx, oracle):
    torch = _torch()
    for pos0, n in ((0, 100000), (12345, 77777), (99990, 10)):
        assert (ctx.gen_bases(7, pos0, n).cpu().numpy() == oracle.gen_bases(7, pos0, n)).all()
        g = ctx.gen_genome(0xB2000003, 100000, 7, 40, 12, pos0, n).cpu().numpy()
        assert (g == oracle.gen_genome(0xB2000003, 100000, 7, 40, 12, pos0, n)).all()
    # misaligned output pointer
    buf = torch.zeros(5000 + 16, dtype=torch.uint8, device="cuda:0")
    ctx.gen_bases(9, 3, 5000, out=buf[5:])
    assert (buf[5:5005].cpu().numpy() == oracle.gen_bases(9, 3, 5000)).all()
    r = ctx.gen_reads(0xB2000004, 50000, 150, 200, 17, 300).cpu().numpy()
    assert (r == oracle.gen_reads(0xB2000004, 50000, 150, 200, 17, 300)).all()


def test_dense_kat_strings(ctx, kmerlib, golden):
    for case in golden["kat_k3"] + golden["dirty"]:
        k, seq = case["k"], case["seq"].encode("latin-1")
        want = np.zeros(4 ** k + 1, dtype=np.int64)
        for i, c in case["counts"].items():
            want[int(i)] = c
        got = dense_gpu(ctx, kmerlib, to_dev(seq), len(seq), k)
        assert (got == want[1:]).all(), case["seq"]


def test_dense_mt19937_and_config1(ctx, kmerlib, golden):
    g = golden["mt19937_64_1mbp_k3"]
    rng = mt19937_64(g["seed"])
    seq = bytes(b"ACGT"[rng.next() & 3] for _ in range(g["n"]))
    got = dense_gpu(ctx, kmerlib, to_dev(seq), len(seq), 3)
    assert got.tolist() == g["counts"][1:]
    g = golden["config1_k3"]
    data = ctx.gen_bases(g["seed"], 0, g["n"])
    got = dense_gpu(ctx, kmerlib, data, g["n"], 3)
    assert got.tolist() == g["counts"][1:]


@pytest.mark.parametrize("k", list(range(1, 14)))
def test_dense_all_k(ctx, kmerlib, oracle, k):
    n = 1_500_000
    s = dirty(oracle, n, seed=k)
    want, _ = oracle.count_dense(s, k)
    d = to_dev(s)
    assert (dense_gpu(ctx, kmerlib, d, n, k) == want).all()
    if 9 <= k <= 12:  # the partition path (auto only picks it for >= 2^26 windows)
        assert (dense_gpu(ctx, kmerlib, d, n, k, algo=kmerlib.DENSE_DIRECT) == want).all()
        assert (dense_gpu(ctx, kmerlib, d, n, k, algo=kmerlib.DENSE_PARTITION) == want).all()


@pytest.mark.parametrize("k", [3, 8, 9, 10, 11, 12, 16])
def test_dense_edges(ctx, kmerlib, oracle, k):
    torch = _torch()
    rng = np.random.default_rng(k)
    algos = [kmerlib.DENSE_DIRECT] + ([kmerlib.DENSE_PARTITION] if 9 <= k <= 12 else [])
    if k == 16:
        # 16 GiB table: only the code path for tiny inputs against a sparse oracle
        s = oracle.gen_bases(1, 0, 3000)
        table = torch.zeros(1 << 32, dtype=torch.int32, device="cuda:0")
        ctx.count_dense_range(to_dev(s), s.size, 0, s.size, 16, table)
        keys, counts, _ = oracle.count_sparse(s, 16)
        idx = torch.from_numpy(keys.astype(np.int64)).cuda()
        assert (table[idx].cpu().numpy().view(np.uint32) == counts).all()
        assert int(table.sum().item()) == int(counts.sum())
        del table
        torch.cuda.empty_cache()
        return
    for n in (0, 1, k - 1, k, k + 1, 15, 16, 17, 511, 512, 513, 1023, 4097, 70001):
        s = rng.choice(list(b"ACGTN"), size=n, p=[.24, .24, .24, .24, .04]).astype(np.uint8) if n else np.zeros(0, np.uint8)
        want, _ = oracle.count_dense(s, k) if n else (np.zeros(4 ** k, np.uint32), 0)
        for algo in algos:
            for off in (0, 1, 7, 15):
                d = _torch().zeros(n + 128, dtype=_torch().uint8, device="cuda:0")
                if n:
                    d[off: off + n] = _torch().from_numpy(s)
                got = dense_gpu(ctx, kmerlib, d, n, k, algo=algo, offset=off)
                assert (got == want).all(), (n, algo, off)


@pytest.mark.parametrize("k,algo", [(5, 1), (12, 1), (12, 2), (9, 2), (10, 2), (11, 2)])
def test_dense_range_additivity(ctx, kmerlib, oracle, k, algo):
    n = 900_001
    s = dirty(oracle, n, seed=40 + k)
    want, _ = oracle.count_dense(s, k)
    nwin = n - k + 1
    cuts = [0, 1, 5, 511, 512, 513, 300_000, 300_007, 650_000, nwin - 1, nwin]
    ranges = list(zip(cuts[:-1], cuts[1:]))
    got = dense_gpu(ctx, kmerlib, to_dev(s), n, k, algo=algo, ranges=ranges)
    assert (got == want).all()
    # shards as the multi-GPU path cuts them: each rank sees only its bytes + halo
    torch = _torch()
    table = torch.zeros(4 ** k, dtype=torch.int32, device="cuda:0")
    for world in (2, 8):
        table.zero_()
        for r in range(world):
            b, e, bb, be = kmerlib.shard_windows(n, k, r, world)
            shard = to_dev(s[bb:be])
            ctx.count_dense_range(shard, be - bb, 0, e - b, k, table, algo=algo)
        torch.cuda.synchronize()
        assert (table.cpu().numpy().view(np.uint32) == want).all(), world


def test_dense_skewed_inputs(ctx, kmerlib, oracle):
    """All-A (one bin, one partition), low-complexity repeats and all-invalid inputs."""
    n = 3_000_000
    for s in (np.full(n, ord("A"), np.uint8), np.frombuffer(b"AC" * (n // 2), np.uint8),
              np.full(n, ord("N"), np.uint8), np.frombuffer(b"ACGTTGCA" * (n // 8), np.uint8)):
        for k, algo in ((4, 0), (12, 1), (12, 2), (9, 2), (10, 2), (11, 2)):
            want, _ = oracle.count_dense(s, k)
            assert (dense_gpu(ctx, kmerlib, to_dev(s), n, k, algo=algo) == want).all()


def test_dense_host_end_to_end(ctx, kmerlib, oracle):
    torch = _torch()
    n = 5_000_000
    s = dirty(oracle, n, seed=3)
    pinned = torch.from_numpy(s.copy()).pin_memory()
    for k in (3, 8, 12):
        want, _ = oracle.count_dense(s, k)
        got = ctx.count_dense_host(pinned, k)
        assert (got == want).all()
        got = ctx.count_dense_host(s, k)  # pageable
        assert (got == want).all()


def test_sync_entry_point_and_errors(ctx, kmerlib, oracle):
    torch = _torch()
    L = kmerlib.lib()
    s = oracle.gen_bases(2, 0, 100000)
    d = to_dev(s)
    table = torch.empty(4 ** 6, dtype=torch.int32, device="cuda:0")
    assert L.kc_count_dense(ctx._h, d.data_ptr(), s.size, 6, table.data_ptr()) == 0
    want, _ = oracle.count_dense(s, 6)
    assert (table.cpu().numpy().view(np.uint32) == want).all()
    assert L.kc_count_dense(ctx._h, d.data_ptr(), s.size, 0, table.data_ptr()) == kmerlib.KC_ERR_INVALID
    assert L.kc_count_dense(ctx._h, d.data_ptr(), s.size, 17, table.data_ptr()) == kmerlib.KC_ERR_INVALID
    assert b"dense k must be" in L.kc_last_error(ctx._h)
    assert L.kc_count_dense_range_async(ctx._h, d.data_ptr(), s.size, 0, s.size, 6, table.data_ptr(), 2, None) == kmerlib.KC_ERR_UNSUPPORTED  # partition path: k = 9..12
    assert ctx.launch_count > 0


# --------------------------------------------------------------------------
def test_per_seq_reference_shape(ctx, kmerlib, oracle, golden):
    """config 1 through the reference-shaped call: data = seq + NUL, offsets = [0, L+1]."""
    torch = _torch()
    g = golden["config1_k3"]
    seq = oracle.gen_bases(g["seed"], 0, g["n"])
    data = np.concatenate([seq, np.zeros(1, np.uint8)])
    offs = torch.tensor([0, g["n"] + 1], dtype=torch.int64, device="cuda:0")
    sums = ctx.count_per_seq(to_dev(data), offs, 1, 3)
    assert sums.shape == (64, 1) and sums[:, 0].cpu().tolist() == g["counts"][1:]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 6, 7, 9])
def test_per_seq_vs_oracle(ctx, kmerlib, oracle, k):
    torch = _torch()
    rng = np.random.default_rng(100 + k)
    lens = [0, 1, k - 1, k, k + 1, 40, 150, 150, 151, 5000, 70000, 3, 200000, 17]
    seqs = [rng.choice(list(b"ACGTNa"), size=n, p=[.24, .24, .24, .24, .03, .01]).astype(np.uint8) for n in lens]
    data = np.concatenate([np.concatenate([s, np.zeros(1, np.uint8)]) for s in seqs])
    offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)
    want, _ = oracle.count_per_seq(data, offs, k)
    got = ctx.count_per_seq(to_dev(data), torch.from_numpy(offs).cuda(), len(seqs), k).cpu().numpy()
    assert (got == want).all()
    # separator content must not matter (kernels.h:133 never reads it): use 'A'
    data2 = data.copy()
    data2[offs[1:] - 1] = ord("A")
    got2 = ctx.count_per_seq(to_dev(data2), torch.from_numpy(offs).cuda(), len(seqs), k).cpu().numpy()
    assert (got2 == want).all()
    # row sums = aggregate dense table of the NUL-separated stream
    dense = dense_gpu(ctx, kmerlib, to_dev(data), data.size, k)
    assert (got.sum(axis=1).astype(np.uint32) == dense).all()


def test_per_seq_many_reads(ctx, kmerlib, oracle):
    torch = _torch()
    reads = oracle.gen_reads(0xB2000004, 100000, 150, 200, 0, 3000)  # '\n' separated, stride 151
    offs = (np.arange(3001) * 151).astype(np.int64)
    want, _ = oracle.count_per_seq(reads, offs, 4)
    got = ctx.count_per_seq(to_dev(reads), torch.from_numpy(offs).cuda(), 3000, 4).cpu().numpy()
    assert (got == want).all()


def test_seqset_to_device_and_count(ctx, kmerlib, oracle, golden):
    torch = _torch()
    case = [c for c in golden["loader"] if c["name"] == "blank_separated" and c["mode"] == 0][0]
    s = kmerlib.SeqSet.from_memory(case["fasta"], 0, 100)
    d_data, d_offs = s.to_device(ctx)
    sums = ctx.count_per_seq(d_data, d_offs, s.num_seqs, 3).cpu().numpy()
    want, _ = oracle.count_per_seq(s.data, s.offsets, 3)
    assert (sums == want).all()
    assert sums[kmerlib.kmer_index("ACG"), 0] == 2 and sums[kmerlib.kmer_index("TTT"), 1] == 2  # ACGTACGGT, TTTT


def test_distance_golden(ctx, kmerlib, oracle, golden):
    torch = _torch()
    for case in golden["distance"]:
        k = case["k"]
        seqs = [s.encode("latin-1") for s in case["seqs"]]
        data = b"".join(s + b"\0" for s in seqs)
        offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)
        d_offs = torch.from_numpy(offs).cuda()
        sums = ctx.count_per_seq(to_dev(data), d_offs, len(seqs), k)
        d = ctx.kmer_distance(sums, d_offs, len(seqs), k).cpu().numpy()
        assert [x.tobytes().hex() for x in d] == case["dist_hex"]


def test_distance_many(ctx, kmerlib, oracle):
    torch = _torch()
    rng = np.random.default_rng(8)
    seqs = [rng.choice(list(b"ACGT"), size=int(rng.integers(50, 3000))).astype(np.uint8) for _ in range(300)]
    data = np.concatenate([np.concatenate([s, np.zeros(1, np.uint8)]) for s in seqs])
    offs = np.cumsum([0] + [len(s) + 1 for s in seqs]).astype(np.int64)
    d_offs = torch.from_numpy(offs).cuda()
    for k in (3, 6):
        sums = ctx.count_per_seq(to_dev(data), d_offs, len(seqs), k)
        got = ctx.kmer_distance(sums, d_offs, len(seqs), k).cpu().numpy()
        want = oracle.distance(sums.cpu().numpy(), offs, k)
        assert got.tobytes() == want.tobytes()


# --------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 5, 16, 17, 18, 21, 31])
@pytest.mark.parametrize("algo", [0, 1])
def test_sparse_vs_oracle(ctx, kmerlib, oracle, k, algo):
    n = 400_000
    s = dirty(oracle, n, seed=60 + k)
    s[100_000:200_000] = s[:100_000]  # repeats so counts exceed 1 at large k
    wk, wc, _ = oracle.count_sparse(s, k)
    sp = ctx.count_sparse(to_dev(s), n, k, algo)
    keys, counts = sp.to_host()
    assert len(sp) == len(wk) and (keys == wk).all() and (counts == wc).all()


def test_sparse_edges(ctx, kmerlib, oracle):
    for algo in (0, 1):
        for s in (b"", b"ACGT", b"N" * 100, b"ACGT" * 20, b"A" * 1000):
            sp = ctx.count_sparse(to_dev(s), len(s), 21, algo)
            wk, wc, _ = oracle.count_sparse(s, 21)
            keys, counts = sp.to_host()
            assert (keys == wk).all() and (counts == wc).all(), (algo, s[:8])
        # tiny capacity hint forces the grow-and-retry path
        s = oracle.gen_bases(4, 0, 200000)
        sp = ctx.count_sparse(to_dev(s), s.size, 21, 0, capacity_hint=16)
        wk, wc, _ = oracle.count_sparse(s, 21)
        keys, counts = sp.to_host()
        assert (keys == wk).all() and (counts == wc).all()


def test_sparse_reads_config4_shape(ctx, kmerlib, oracle):
    """config 4/5 in miniature: 150 bp reads with substitutions, '\\n' separated."""
    reads = oracle.gen_reads(0xB2000004, 200_000, 150, 200, 0, 20_000)
    d = to_dev(reads)
    for k in (21, 31):
        wk, wc, _ = oracle.count_sparse(reads, k)
        for algo in (0, 1):
            keys, counts = ctx.count_sparse(d, reads.size, k, algo).to_host()
            assert (keys == wk).all() and (counts == wc).all()
        assert int(wc.sum()) == 20_000 * (150 - k + 1)


def test_sparse_owner_sharding(ctx, kmerlib, oracle):
    """hash-sharded path in one process: G ranks count their reads, bucket by owner,
    'exchange', merge — the union of the owners equals the global result."""
    torch = _torch()
    G, k, nreads = 4, 21, 8000
    reads = oracle.gen_reads(0xB2000004, 100_000, 150, 200, 0, nreads)
    wk, wc, _ = oracle.count_sparse(reads, k)
    inbox = [[] for _ in range(G)]
    for r in range(G):
        r0, r1 = kmerlib.shard_reads(nreads, r, G)
        shard = to_dev(reads[r0 * 151: r1 * 151])
        sp = ctx.count_sparse(shard, (r1 - r0) * 151, k, 0)
        n = len(sp)
        ok = torch.empty(n, dtype=torch.int64, device="cuda:0")
        oc = torch.empty(n, dtype=torch.int32, device="cuda:0")
        sizes = ctx.sparse_bucket_by_owner(sp.d_keys, sp.d_counts, n, G, ok, oc)
        assert int(sizes.sum()) == n
        start = 0
        for o in range(G):
            m = int(sizes[o])
            keys_o = ok[start:start + m]
            assert all(kmerlib.mix64(int(x)) % G == o for x in keys_o[:50].cpu().tolist())
            inbox[o].append((keys_o.clone(), oc[start:start + m].clone()))
            start += m
    allk, allc = [], []
    for o in range(G):
        kk = torch.cat([a for a, _ in inbox[o]])
        cc = torch.cat([b for _, b in inbox[o]])
        keys, counts = ctx.sparse_merge(kk, cc, kk.numel()).to_host()
        assert (np.diff(keys.astype(np.int64)) > 0).all()
        allk.append(keys)
        allc.append(counts)
    keys = np.concatenate(allk)
    counts = np.concatenate(allc)
    order = np.argsort(keys)
    assert (keys[order] == wk).all() and (counts[order] == wc).all()


# --------------------------------------------------------------------------
def test_config3_full_size(ctx, kmerlib, oracle):
    """BASELINE config 3 at full size (3.1 Gbp, k=12, N runs): the partition path,
    the direct path and a 4-shard decomposition must agree bit for bit, satisfy
    sum(counts) + invalid = L - k + 1, and match the multi-threaded oracle."""
    torch = _torch()
    L, k, seed = 3_100_000_000, 12, 0xB2000003
    data = ctx.gen_genome(seed, L, 1000, 10000, k, 0, L)
    t_part = torch.zeros(4 ** k, dtype=torch.int32, device="cuda:0")
    ctx.count_dense_range(data, L, 0, L, k, t_part, algo=kmerlib.DENSE_PARTITION)
    t_dir = torch.zeros(4 ** k, dtype=torch.int32, device="cuda:0")
    ctx.count_dense_range(data, L, 0, L, k, t_dir, algo=kmerlib.DENSE_DIRECT)
    t_sh = torch.zeros(4 ** k, dtype=torch.int32, device="cuda:0")
    for r in range(4):
        b, e, bb, be = kmerlib.shard_windows(L, k, r, 4)
        ctx.count_dense_range(data[bb:be], be - bb, 0, e - b, k, t_sh)
    torch.cuda.synchronize()
    assert torch.equal(t_part, t_dir) and torch.equal(t_part, t_sh)
    total = int(t_part.to(torch.int64).sum().item())
    host = data.cpu().numpy()
    del data
    threads = min(os.cpu_count() or 8, 48)
    want, inv = oracle.count_dense(host, k, threads=threads)
    assert total + inv == L - k + 1
    assert (t_part.cpu().numpy().view(np.uint32) == want).all()


def test_all_256_byte_values(ctx, kmerlib, oracle):
    """Only the four upper-case letters may pass the SIMD validity test of the decoder."""
    rng = np.random.default_rng(256)
    every = np.arange(256, dtype=np.uint8)
    s = np.concatenate([every, rng.permutation(every), np.repeat(every, 3), rng.integers(0, 256, 100000).astype(np.uint8)])
    for k in (1, 2, 3):
        want, _ = oracle.count_dense(s, k)
        assert (dense_gpu(ctx, kmerlib, to_dev(s), s.size, k) == want).all()
    want1, _ = oracle.count_dense(every, 1)
    assert want1.tolist() == [1, 1, 1, 1]
    # the same bytes through the k=12 partition path, embedded in valid sequence
    base = oracle.gen_bases(3, 0, 1 << 22)
    base[1000:1000 + s.size:97] = s[: len(base[1000:1000 + s.size:97])]
    want, _ = oracle.count_dense(base, 12)
    assert (dense_gpu(ctx, kmerlib, to_dev(base), base.size, 12, algo=kmerlib.DENSE_PARTITION) == want).all()


def test_dense_k8_smem16(ctx, kmerlib, oracle):
    """k=8 shared-memory path (16-bit fields with bounded spill): random, dirty and
    maximally skewed inputs, unaligned pointer, range split."""
    n = 9_000_001
    for s in (dirty(oracle, n, seed=88), np.full(n, ord("A"), np.uint8), np.frombuffer(b"ACGTTGCAAC" * (n // 10), np.uint8)):
        m = s.size
        want, _ = oracle.count_dense(s, 8)
        d = to_dev(s)
        assert (dense_gpu(ctx, kmerlib, d, m, 8) == want).all()
        assert (dense_gpu(ctx, kmerlib, d, m, 8, algo=kmerlib.DENSE_DIRECT) == want).all()
    s = dirty(oracle, n, seed=89)
    want, _ = oracle.count_dense(s, 8)
    d = _torch().zeros(n + 128, dtype=_torch().uint8, device="cuda:0")
    d[5: 5 + n] = _torch().from_numpy(s)
    assert (dense_gpu(ctx, kmerlib, d, n, 8, offset=5) == want).all()
    nwin = n - 7
    assert (dense_gpu(ctx, kmerlib, to_dev(s), n, 8, ranges=[(0, 4_500_003), (4_500_003, nwin)]) == want).all()


def test_sparse_deep_coverage_hash_equals_sort(ctx, kmerlib, oracle):
    """config 4 shape at 1/20 scale (5 M reads, 750 Mbp): too big for the CPU oracle's sort
    in test time, so check the two independent GPU algorithms against each other and
    against sum(counts) = number of windows (reads hold only ACGT)."""
    nreads, k = 5_000_000, 21
    reads = ctx.gen_reads(0xB2000004, 25_000_000, 150, 200, 0, nreads)
    a = ctx.count_sparse(reads, nreads * 151, k, kmerlib.SPARSE_HASH)
    b = ctx.count_sparse(reads, nreads * 151, k, kmerlib.SPARSE_SORT)
    ka, ca = a.to_host()
    kb, cb = b.to_host()
    assert len(a) == len(b) and (ka == kb).all() and (ca == cb).all()
    assert int(ca.astype(np.int64).sum()) == nreads * (150 - k + 1)
    assert (np.diff(ka.astype(np.int64)) > 0).all() and int(ka.max()) < (1 << 42)
    # a slice of it against the oracle
    part = reads[: 20_000 * 151].cpu().numpy()
    wk, wc, _ = oracle.count_sparse(part, k)
    kp, cp = ctx.count_sparse(reads, 20_000 * 151, k, kmerlib.SPARSE_SORT).to_host()
    assert (kp == wk).all() and (cp == wc).all()


def test_sparse_unsorted_flag(ctx, kmerlib, oracle):
    s = dirty(oracle, 300_000, seed=77)
    wk, wc, _ = oracle.count_sparse(s, 21)
    keys, counts = ctx.count_sparse(to_dev(s), s.size, 21, kmerlib.SPARSE_HASH | kmerlib.SPARSE_UNSORTED).to_host()
    order = np.argsort(keys)
    assert (keys[order] == wk).all() and (counts[order] == wc).all()


def test_config2_full_size(ctx, kmerlib, oracle):
    """BASELINE configs[1] at its real size (100 Mbp, k = 8, the bench's seed): the library default (16-bit
    shared-memory bins with the per-CTA checksum) against the CPU oracle, bin by bin"""
    n, seed = 100_000_000, 0xB2000002
    data = ctx.gen_genome(seed, n, 0, 0, 8, 0, n)
    want, _ = oracle.count_dense(oracle.gen_genome(seed, n, 0, 0, 8, 0, n), 8, threads=8)
    assert (dense_gpu(ctx, kmerlib, data, n, 8) == want).all()
    assert ctx.dense_fingerprint(_table_of(ctx, kmerlib, data, n, 8), 8) == ctx.window_fingerprint(data, n, 8)


def _table_of(ctx, kmerlib, d, n, k):
    t = _torch().zeros(kmerlib.num_kmers(k), dtype=_torch().int32, device="cuda:0")
    ctx.count_dense_range(d, n, 0, n, k, t)
    _torch().cuda.synchronize()
    return t


def test_config4_full_scale_self_check(ctx, kmerlib):
    """BASELINE configs[3] at FULL scale on one GPU (100 M reads x 150 bp, k = 21, the bench's generator): no CPU oracle
    goes there, so the count is checked by the multiset identity of csrc/check.cu — the window fingerprint of the input
    (one streaming scan, pinned to the oracle in test_fingerprint_self_checks) must equal the fingerprint of the result,
    the sum of counts the number of valid windows, the keys must ascend strictly — and by the number of distinct k-mers,
    which every algorithm and every GPU count (1, 2, 4, 8) of round 2 agreed on."""
    nreads, k = 100_000_000, 21
    reads = ctx.gen_reads(0xB2000004, 500_000_000, 150, 200, 0, nreads)
    nb = nreads * 151
    fin = ctx.window_fingerprint(reads, nb, k)
    assert fin[1] == nreads * (150 - k + 1)
    sp = ctx.count_sparse(reads, nb, k, kmerlib.SPARSE_AUTO)
    assert ctx.sparse_fingerprint(sp) == fin + (0,)
    assert len(sp) == 1_774_844_036
    sp.close()


def test_config5_quarter_scale_self_check(ctx, kmerlib):
    """BASELINE configs[4] at 1/4 scale (50 M reads of 200 M, 1 Gbp genome, k = 31: 64-bit records on both partition
    levels), radix and hash table: same self-check, same distinct count"""
    nreads, k = 50_000_000, 31
    ctx.release_memory()  # the slabs of the config-4 test above (60 GB) would leave the hash table too little room
    _torch().cuda.empty_cache()
    reads = ctx.gen_reads(0xB2000005, 1_000_000_000, 150, 200, 0, nreads)
    nb = nreads * 151
    fin = ctx.window_fingerprint(reads, nb, k)
    assert fin[1] == nreads * (150 - k + 1)
    for algo in (kmerlib.SPARSE_AUTO, kmerlib.SPARSE_HASH):
        sp = ctx.count_sparse(reads, nb, k, algo)
        assert ctx.sparse_fingerprint(sp) == fin + (0,), algo
        assert len(sp) == 1_854_369_236
        sp.close()

"""The oracle (oracle/kmer_oracle.c) against the reference: golden vectors that
were produced by the reference's own code (tests/golden/make_golden.py), and —
where oracle/_ref is present — the reference's code live.  CPU only."""
import numpy as np
import pytest

from mt64 import mt19937_64


def _expect(counts_dict, k):
    v = np.zeros(4 ** k + 1, dtype=np.int64)
    for i, c in counts_dict.items():
        v[int(i)] = c
    return v


def test_permutation_order(oracle, golden):
    for k in (3, 4):
        assert oracle.permutation("ACGT", k) == golden["permutation"][str(k)]
    for k in (3, 4, 5, 6):
        joined = "".join(oracle.permutation("ACGT", k)).encode()
        assert "%016x" % oracle.fnv1a64(np.frombuffer(joined, dtype=np.uint8)) == golden["permutation_fnv"][str(k)]
    p = oracle.permutation("ACGT", 3)
    assert p[1] == "CAA" and p[16] == "AAC"  # little-endian in the string (SURVEY §0 rule 5)
    assert oracle.kmer_index("ACG") == 36 and oracle.kmer_index("CGT") == 57 and oracle.kmer_index("TTT") == 63
    assert oracle.kmer_index("ACN") is None


def test_kat_strings(oracle, golden):
    for case in golden["kat_k3"] + golden["dirty"]:
        k, seq = case["k"], case["seq"].encode("latin-1")
        want = _expect(case["counts"], k)
        naive = oracle.count_all_naive(seq, k)
        assert (naive == want).all(), case["seq"]
        table, inv = oracle.count_dense(seq, k)
        assert (table == want[1:]).all() and inv == want[0], case["seq"]


def test_mt19937_vector(oracle, golden):
    g = golden["mt19937_64_1mbp_k3"]
    rng = mt19937_64(g["seed"])
    seq = bytes(b"ACGT"[rng.next() & 3] for _ in range(g["n"]))
    assert seq[:32].decode() == g["head"]
    table, inv = oracle.count_dense(seq, 3)
    assert inv == g["counts"][0] and table.tolist() == g["counts"][1:]
    assert table[:4].tolist() == [15729, 15702, 15596, 15754]


def test_config1_vector(oracle, golden):
    g = golden["config1_k3"]
    seq = oracle.gen_bases(g["seed"], 0, g["n"])
    assert seq[:32].tobytes().decode() == g["head"]
    table, inv = oracle.count_dense(seq, 3)
    assert inv == 0 and table.tolist() == g["counts"][1:]
    for k in (4, 5, 6):
        gg = golden["config1_prefix200k_k%d" % k]
        t, inv = oracle.count_dense(seq[: gg["n"]], k)
        assert inv == gg["invalid"]
        assert "%016x" % oracle.fnv1a64(t.astype("<i4")) == gg["table_fnv"]


@pytest.mark.parametrize("k", [3, 4, 5, 6])
def test_live_reference(oracle, k):
    """Same inputs through the reference's permutationsCountAll, when built here."""
    ref = oracle.ref(k)
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(k)
    assert oracle.permutation("ACGT", k) == ref.permutation()
    for n in (0, 1, k - 1, k, 100, 5000):
        s = rng.choice(list(b"ACGTNacgt\r|"), size=n, p=[.22, .22, .22, .22, .04, .02, .02, .01, .01, .01, .01]).astype(np.uint8)
        want = ref.count_all(s)
        table, inv = oracle.count_dense(s, k)
        assert (table == want[1:].astype(np.uint32)).all() and inv == want[0]


def test_rolling_vs_naive_all_k(oracle):
    rng = np.random.default_rng(7)
    s = rng.choice(list(b"ACGTN"), size=3000, p=[.24, .24, .24, .24, .04]).astype(np.uint8)
    for k in range(1, 11):
        naive = oracle.count_all_naive(s, k)
        table, inv = oracle.count_dense(s, k)
        assert (table == naive[1:]).all() and inv == naive[0]
        assert int(table.sum()) + inv == max(0, s.size - k + 1)


def test_range_and_threads(oracle):
    s = oracle.gen_genome(99, 200000, 5, 50, 12, 0, 200000)
    for k in (5, 12):
        full, inv = oracle.count_dense(s, k)
        acc = np.zeros_like(full)
        tot_inv = 0
        nwin = s.size - k + 1
        cuts = [0, 1, 777, 65536, 65537, 150001, nwin]
        for a, b in zip(cuts[:-1], cuts[1:]):
            _, i = oracle.count_dense_range(s, k, a, b, acc)
            tot_inv += i
        assert (acc == full).all() and tot_inv == inv
        mt, inv_mt = oracle.count_dense(s, k, threads=5)
        assert (mt == full).all() and inv_mt == inv


def test_marginalisation(oracle):
    """count(k) sums to count(k-1) over the LAST base, up to the final window."""
    s = oracle.gen_bases(5, 0, 50000)
    for k in (4, 9):
        hi, _ = oracle.count_dense(s, k)
        lo, _ = oracle.count_dense(s, k - 1)
        marg = hi.reshape(4, -1).sum(axis=0)  # drop the most significant digit = last base
        last = oracle.kmer_index(s[-(k - 1):].tobytes().decode())
        lo2 = lo.copy()
        lo2[last] -= 1
        assert (marg == lo2).all()


def test_per_seq(oracle, golden):
    seqs = [b"ACGTACGTAC", b"ACGTTTTTAC", b"GG", b"", b"ACGTACGTACGTNNACGT"]
    data = b"".join(s + b"\0" for s in seqs)
    offs = np.cumsum([0] + [len(s) + 1 for s in seqs])
    for k in (1, 3, 5):
        sums, inv = oracle.count_per_seq(data, offs, k)
        for e, s in enumerate(seqs):
            want = oracle.count_all_naive(s, k)
            assert (sums[:, e] == want[1:]).all() and inv[e] == want[0]
        dense, _ = oracle.count_dense(data, k)
        assert (sums.sum(axis=1).astype(np.uint32) == dense).all()


def test_sparse_matches_dense(oracle):
    s = oracle.gen_genome(3, 40000, 3, 20, 9, 0, 40000)
    for k in (3, 9, 13):
        keys, counts, inv = oracle.count_sparse(s, k)
        table, inv2 = oracle.count_dense(s, k)
        nz = np.nonzero(table)[0]
        assert inv == inv2 and (keys == nz.astype(np.uint64)).all() and (counts == table[nz]).all()
    keys, counts, _ = oracle.count_sparse(b"ACGT" * 20, 31)
    assert len(keys) == 4 and counts.sum() == 80 - 31 + 1
    assert keys.max() < (1 << 62)


def test_distance_golden(oracle, golden):
    for case in golden["distance"]:
        k = case["k"]
        seqs = [s.encode("latin-1") for s in case["seqs"]]
        data = b"".join(s + b"\0" for s in seqs)
        offs = np.cumsum([0] + [len(s) + 1 for s in seqs])
        sums, _ = oracle.count_per_seq(data, offs, k)
        d = oracle.distance(sums, offs, k)
        assert [x.tobytes().hex() for x in d] == case["dist_hex"]
    for i, j, n, want in golden["triangular_index"]:
        assert oracle.triangular_index(i, j, n) == want


def test_loader_golden(oracle, golden):
    for case in golden["loader"]:
        r = oracle.import_seqs_mem(case["fasta"], case["mode"], 100)
        assert r["num_seqs"] == case["num_seqs"], case["name"]
        assert r["ids"] == case["ids"], case["name"]
        ref_offs = case["offsets"]
        # the engine always emits the terminal offset; the reference only when the
        # last record ends at EOF / MAX_SEQS (SURVEY §8a row a5 rule 5)
        assert r["offsets"].tolist()[: len(ref_offs)] == ref_offs, case["name"]
        assert len(r["offsets"]) == case["num_seqs"] + 1
        assert r["data"].decode("latin-1") == case["data"], case["name"]
        want_seqs = [s[:-1].replace("|", "\0") for s in case["seqs"]]
        got = [r["data"][r["offsets"][i]: r["offsets"][i + 1] - 1].decode("latin-1") for i in range(r["num_seqs"])]
        assert got == want_seqs, case["name"]
    g = golden["loader_max_seqs"]
    many = "".join(">r%d\nACGT\nGGCC\nTTAA\n\n" % i for i in range(g["records"]))
    r = oracle.import_seqs_mem(many, 0, g["max_seqs"])
    assert r["num_seqs"] == g["num_seqs"]
    assert r["offsets"].tolist()[-3:] == g["offsets_tail"]
    last = r["data"][r["offsets"][-2]: r["offsets"][-1] - 1].decode()
    assert last + "|" == g["last_seq"]


def test_generators_are_positional(oracle):
    whole = oracle.gen_genome(0xB2000003, 100000, 7, 40, 12, 0, 100000)
    part = oracle.gen_genome(0xB2000003, 100000, 7, 40, 12, 33333, 4444)
    assert (whole[33333:33333 + 4444] == part).all()
    assert (whole == ord("N")).sum() > 0
    reads = oracle.gen_reads(0xB2000004, 5000, 150, 200, 0, 64)
    part = oracle.gen_reads(0xB2000004, 5000, 150, 200, 10, 5)
    assert (reads.reshape(64, 151)[10:15].ravel() == part).all()
    assert (reads.reshape(64, 151)[:, 150] == ord("\n")).all()
    assert set(np.unique(reads)) <= set(b"ACGT\n")

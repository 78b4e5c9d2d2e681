"""Host side of the C ABI on a CPU-only box: the library loads, exports every
symbol include/kmer_b200.h declares, the ctx-less entry points (k selection,
FASTA load, dumps) match the reference goldens, and compute entry points fail
loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_exports_match_header(kmerlib):
    declared = kmerlib.declared_symbols()
    assert len(declared) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", kmerlib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported, sorted(set(declared) - exported)
    # nothing but the C ABI leaks out
    assert all(s.startswith("kc_") for s in exported), sorted(s for s in exported if not s.startswith("kc_"))
    assert kmerlib.lib().kc_version() >= 100


def test_no_oracle_in_product(kmerlib):
    """The product library must not link or name the test-side oracle."""
    out = subprocess.run(["ldd", kmerlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    blob = open(kmerlib.LIB_PATH, "rb").read()
    assert b"liboracle" not in blob and b"or_count_dense" not in blob


def test_num_kmers_and_index(kmerlib, golden):
    assert kmerlib.num_kmers(3) == 64 and kmerlib.num_kmers(12) == 1 << 24 and kmerlib.num_kmers(31) == 1 << 62
    assert kmerlib.num_kmers(0) == 0 and kmerlib.num_kmers(32) == 0
    assert kmerlib.kmer_index("ACG") == 36 and kmerlib.kmer_index("CGT") == 57 and kmerlib.kmer_index("TTT") == 63
    for i in (0, 1, 16, 63):
        assert kmerlib.kmer_index(kmerlib.kmer_string(i, 3)) == i
    with pytest.raises(kmerlib.KmerError):
        kmerlib.kmer_index("ACN")
    assert kmerlib.kmer_string((1 << 62) - 1, 31) == "T" * 31


def test_permutation_matches_reference(kmerlib, golden, oracle):
    for k in (3, 4):
        assert kmerlib.permutation("ACGT", k) == golden["permutation"][str(k)]
    for k in (1, 2, 5, 6):
        assert kmerlib.permutation("ACGT", k) == oracle.permutation("ACGT", k)
    for k in (3, 4, 5, 6):
        joined = "".join(kmerlib.permutation("ACGT", k)).encode()
        assert "%016x" % oracle.fnv1a64(np.frombuffer(joined, dtype=np.uint8)) == golden["permutation_fnv"][str(k)]
    assert kmerlib.permutation("01", 3) == ["000", "100", "010", "110", "001", "101", "011", "111"]
    perms = kmerlib.permutation("ACGT", 4)
    assert all(kmerlib.kmer_index(p) == i for i, p in enumerate(perms))


def test_import_seqs_golden(kmerlib, golden, tmp_path):
    for case in golden["loader"]:
        for via_file in (False, True):
            if via_file:
                p = tmp_path / ("%s_%d.fasta" % (case["name"], case["mode"]))
                p.write_bytes(case["fasta"].encode("latin-1"))
                s = kmerlib.SeqSet.from_file(str(p), case["mode"], 100)
            else:
                s = kmerlib.SeqSet.from_memory(case["fasta"], case["mode"], 100)
            assert s.num_seqs == case["num_seqs"], case["name"]
            assert s.ids == case["ids"], case["name"]
            offs = s.offsets.tolist()
            assert offs[: len(case["offsets"])] == case["offsets"], case["name"]
            assert len(offs) == case["num_seqs"] + 1 and offs[-1] == s.nbytes
            assert s.data.decode("latin-1") == case["data"], case["name"]
            s.close()


def test_import_seqs_max_seqs(kmerlib, golden, oracle):
    g = golden["loader_max_seqs"]
    many = "".join(">r%d\nACGT\nGGCC\nTTAA\n\n" % i for i in range(g["records"]))
    s = kmerlib.SeqSet.from_memory(many, 0, g["max_seqs"])
    assert s.num_seqs == g["num_seqs"] and s.offsets.tolist()[-3:] == g["offsets_tail"]
    unlimited = kmerlib.SeqSet.from_memory(many, 0, 0)
    assert unlimited.num_seqs == g["records"]
    r = oracle.import_seqs_mem(many, 0, 0)
    assert unlimited.data == r["data"] and unlimited.offsets.tolist() == r["offsets"].tolist()


def test_import_seqs_random_vs_oracle(kmerlib, oracle):
    rng = np.random.default_rng(11)
    pieces = [b">h\n", b"ACGT\n", b"NNAC\n", b"\n", b"\r\n", b"acgt\n", b">x y\n", b"GG|TT\n", b"T", b"\n\n", b"CCC\r\n"]
    for trial in range(200):
        text = b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=int(rng.integers(0, 25))))
        for mode in (0, 1):
            for mx in (0, 2):
                s = kmerlib.SeqSet.from_memory(text, mode, mx)
                r = oracle.import_seqs_mem(text, mode, mx)
                assert s.num_seqs == r["num_seqs"] and s.ids == r["ids"], (text, mode, mx)
                assert s.data == r["data"] and s.offsets.tolist() == r["offsets"].tolist(), (text, mode, mx)


def test_import_seqs_threads_equal_serial(kmerlib, oracle, golden):
    """kc_import_seqs_mem_threads (f2: ingest at speed) == the serial loader == the oracle, for
    every thread count: chunks start in all three parser states, records span chunks, files end
    inside a record / after a blank line / on a header."""
    rng = np.random.default_rng(23)
    pieces = [b">h\n", b"ACGT\n", b"NNAC\n", b"\n", b"\r\n", b"acgt\n", b">x y\n", b"GG|TT\n", b"T", b"\n\n", b"CCC\r\n",
              b"ACGTACGTACGTACGTACGT\n", b">\n", b"|\n"]
    texts = [case["fasta"].encode("latin-1") for case in golden["loader"]]
    for trial in range(120):
        texts.append(b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=int(rng.integers(0, 60)))))
    for text in texts:
        for mode in (0, 1):
            want = kmerlib.SeqSet.from_memory(text, mode, 0)  # serial path (input below 32 MiB)
            r = oracle.import_seqs_mem(text, mode, 0)
            assert want.data == r["data"] and want.offsets.tolist() == r["offsets"].tolist()
            for nt in (1, 2, 3, 5, 9):
                s = kmerlib.SeqSet.from_memory_threads(text, mode, nt)
                assert s.num_seqs == want.num_seqs and s.ids == want.ids, (text, mode, nt)
                assert s.data == want.data and s.offsets.tolist() == want.offsets.tolist(), (text, mode, nt)
                s.close()
            want.close()


def test_import_seqs_large_file_takes_threaded_path(kmerlib, oracle, tmp_path):
    """40 MiB FASTA (70-column lines, blank-line separated records) through kc_import_seqs: the
    threaded parser is picked by size; result equals the oracle's serial restatement."""
    rng = np.random.default_rng(3)
    recs = []
    for i in range(40):
        seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, 1 << 20)].tobytes()
        lines = b"\n".join(seq[j:j + 70] for j in range(0, len(seq), 70))
        recs.append(b">chr%d test\n" % i + lines + b"\n\n")
    text = b"".join(recs)
    assert len(text) > (32 << 20)
    p = tmp_path / "big.fasta"
    p.write_bytes(text)
    for mode in (0, 1):
        s = kmerlib.SeqSet.from_file(str(p), mode, 0)
        r = oracle.import_seqs_mem(text, mode, 0)
        assert s.num_seqs == r["num_seqs"] == 40 and s.ids == r["ids"]
        assert s.data == r["data"] and s.offsets.tolist() == r["offsets"].tolist()
        s.close()


def test_import_seqs_live_reference(kmerlib, oracle, tmp_path):
    ref = oracle.ref(3)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    pieces = [b">h\n", b"ACGT\n", b"NNAC\n", b"\n", b"\r\n", b"acgt\n", b">x y\n", b"GG|TT\n", b"T", b"\n\n"]
    for trial in range(60):
        text = b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=int(rng.integers(0, 20))))
        p = tmp_path / "t.fasta"
        p.write_bytes(text)
        for mode in (0, 1):
            r = ref.import_seqs(str(p), mode)
            s = kmerlib.SeqSet.from_file(str(p), mode, 100)
            assert s.num_seqs == r["num_seqs"] and s.ids == r["ids"], (text, mode)
            assert s.data == r["data"], (text, mode)
            assert s.offsets.tolist()[: len(r["offsets"])] == r["offsets"].tolist(), (text, mode)


def test_import_missing_file(kmerlib):
    with pytest.raises(kmerlib.KmerError) as e:
        kmerlib.SeqSet.from_file("/nonexistent/all_seqs.fasta")
    assert e.value.code == kmerlib.KC_ERR_IO and "Error opening" in str(e.value)  # text of main.cu:478


def test_dump_counts_format(kmerlib, oracle, tmp_path):
    seqs = [b"ACGTACGTAC", b"ACGTTTTTAC", b"GGGGGGGG"]
    data = b"".join(s + b"\0" for s in seqs)
    offs = np.cumsum([0] + [len(s) + 1 for s in seqs])
    sums, _ = oracle.count_per_seq(data, offs, 3)
    p = tmp_path / "sums.txt"
    kmerlib.dump_counts(str(p), sums, 3, 3)
    text = p.read_bytes()
    assert text == oracle.dump_counts(sums, 3, 3)
    lines = text.decode().split("\n")
    assert lines[0] == "Sums:" and lines[1] == "0: 0,\t0,\t0,\t" and lines[37] == "36: 2,\t1,\t0,\t"
    assert text.endswith(b"\n\n") and len(lines) == 64 + 3
    d = np.array([0.625, 1.0, 0.0], dtype=np.float32)
    kmerlib.dump_distances(str(p), d)
    assert p.read_text() == "0.625000\n1.000000\n0.000000\n"


def test_triangular_index(kmerlib, golden):
    for i, j, n, want in golden["triangular_index"]:
        assert kmerlib.triangular_index(i, j, n) == want


def test_mix64_matches_oracle(kmerlib, oracle):
    for x in (0, 1, 0xDEADBEEF, (1 << 62) - 1, 12345678901234567):
        assert kmerlib.mix64(x) == oracle.mix64(x)


def test_shard_windows_cover(kmerlib):
    for nbytes, k in ((1000, 12), (11, 12), (12, 12), (3_100_000_000, 12), (17, 3)):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                b, e, bb, be = kmerlib.shard_windows(nbytes, k, r, world)
                assert b == prev and e >= b
                if e > b:
                    assert bb == b and be == e + k - 1 <= nbytes
                prev = e
            assert prev == max(nbytes - k + 1, 0)


def test_no_gpu_fails_loudly(kmerlib):
    """Without a device the product refuses to work — it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kmerlib.KmerError) as e:
        kmerlib.Context(0)
    assert e.value.code == kmerlib.KC_ERR_CUDA and "no CPU fallback" in str(e.value)
    # a NULL ctx is rejected by every compute entry point
    L = kmerlib.lib()
    assert L.kc_count_dense(None, None, 0, 3, None) == kmerlib.KC_ERR_INVALID
    assert L.kc_count_per_seq(None, None, None, 0, 3, None) == kmerlib.KC_ERR_INVALID
    assert L.kc_count_sparse(None, None, 0, 21, 0, 0, None) == kmerlib.KC_ERR_INVALID
    assert L.kc_count_dense_host_packed(None, None, 0, 3, None, 0) == kmerlib.KC_ERR_INVALID


def _store_reference(data):
    """the 2-bit store of include/kmer_b200.h in numpy: packed bytes (first base on top, main.cu:83) + validity bitmap"""
    n = data.size
    code = np.full(256, -1, dtype=np.int64)
    for i, ch in enumerate(b"ACGT"):
        code[ch] = i
    c = code[data]
    bad = c < 0
    q = np.concatenate([np.where(bad, 0, c).astype(np.uint8), np.zeros((-n) % 4, np.uint8)]).reshape(-1, 4)
    packed = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
    bits = np.concatenate([bad, np.ones((-n) % 32, dtype=bool)]).reshape(-1, 32)
    mask = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    return packed.astype(np.uint8), mask


def test_pack_2bit_host_matches_store_layout(kmerlib):
    """kc_pack_2bit_host (format conversion on the host cores, no ctx): every byte value, lengths around the
    32-base word and the 2^20-base item edges, AVX-512 / AVX2 / scalar bodies, 1..5 threads"""
    rng = np.random.default_rng(11)
    assert kmerlib.lib().kc_host_pack_simd() in (0, 1, 2)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for n in (1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 4096 + 17, (1 << 20) - 1, (1 << 20) + 33, 3 * (1 << 20) + 5):
        data = acgt[rng.integers(0, 4, n)].copy()
        dirty = rng.random(n) < 0.02
        data[dirty] = rng.integers(0, 256, int(dirty.sum())).astype(np.uint8)
        if n >= 256:
            data[:256] = np.arange(256, dtype=np.uint8)
        wp, wm = _store_reference(data)
        for nthreads, body in ((1, 0), (5, 0), (1, 1), (3, 1), (1, 2), (2, 2)):
            p, m = kmerlib.pack_2bit_host(data, nthreads, body)
            assert (p == wp).all() and (m == wm).all(), (n, nthreads, body)
    p, m = kmerlib.pack_2bit_host(np.zeros(0, dtype=np.uint8))
    assert p.size == 0 and m.size == 0


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/kmer_b200.h must compile as C99 (and as C++11) without any CUDA or torch header"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hdr.c"
    src.write_text('#include "kmer_b200.h"\nint main(void) { return kc_version() > 0 ? 0 : 1; }\n')
    inc = os.path.join(root, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + inc, "-fsyntax-only", str(src)],
                ["g++", "-std=c++11", "-Wall", "-Werror", "-I" + inc, "-fsyntax-only", "-x", "c++", str(src)]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

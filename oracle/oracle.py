"""ctypes bindings for the TEST-SIDE oracle (oracle/liboracle.so) and, when it
was built, the reference's own host code (oracle/_ref/libref_k<K>.so).

*** TEST INFRASTRUCTURE ONLY ***  Import this from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never from the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = {}

u8p = C.POINTER(C.c_ubyte)


def build(ref=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    targets = ["all"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.or_num_kmers.restype = C.c_uint64
        L.or_num_kmers.argtypes = [C.c_int]
        L.or_kmer_index.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
        L.or_count_all_naive.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_void_p]
        L.or_count_dense_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                           C.c_void_p, C.POINTER(C.c_uint64)]
        L.or_count_dense.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
        L.or_count_dense_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p,
                                        C.POINTER(C.c_uint64), C.c_int]
        L.or_count_per_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.or_count_sparse.restype = C.c_int64
        L.or_count_sparse.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.or_free.argtypes = [C.c_void_p]
        L.or_triangular_index.restype = C.c_int64
        L.or_triangular_index.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.or_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.or_import_seqs_mem.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_long, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        L.or_dump_counts.restype = C.c_void_p
        L.or_dump_counts.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.or_sm64.restype = C.c_uint64
        L.or_sm64.argtypes = [C.c_uint64]
        L.or_mix64.restype = C.c_uint64
        L.or_mix64.argtypes = [C.c_uint64]
        L.or_gen_bases.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.or_gen_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64,
                                    C.c_uint64, C.c_void_p]
        L.or_gen_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                   C.c_void_p]
        L.or_fnv1a64.restype = C.c_uint64
        L.or_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        _LIB = L
    return _LIB


def _bytes_arr(data):
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    if isinstance(data, str):
        return np.frombuffer(data.encode("latin-1"), dtype=np.uint8)
    return np.ascontiguousarray(data, dtype=np.uint8)


def num_kmers(k):
    return int(lib().or_num_kmers(k))


def permutation(alphabet, k):
    n = len(alphabet) ** k
    bufs = [C.create_string_buffer(k + 1) for _ in range(n)]
    arr = (C.c_char_p * n)(*[C.cast(b, C.c_char_p) for b in bufs])
    lib().or_permutation.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
    lib().or_permutation(alphabet.encode(), k, arr)
    return [b.value.decode() for b in bufs]


def kmer_index(s, k=None):
    k = k or len(s)
    out = C.c_uint64()
    rc = lib().or_kmer_index(s.encode("latin-1"), k, C.byref(out))
    return None if rc else int(out.value)


def count_all_naive(seq, k):
    """Reference-CPU-shaped result: [0] invalid bucket, [idx+1] counts."""
    a = _bytes_arr(seq)
    out = np.zeros(num_kmers(k) + 1, dtype=np.int64)
    lib().or_count_all_naive(a.tobytes(), a.size, k, out.ctypes.data)
    return out


def count_dense(data, k, threads=1):
    a = _bytes_arr(data)
    table = np.zeros(num_kmers(k), dtype=np.uint32)
    inv = C.c_uint64(0)
    if threads > 1:
        rc = lib().or_count_dense_mt(a.ctypes.data, a.size, k, table.ctypes.data, C.byref(inv), threads)
        assert rc == 0
    else:
        lib().or_count_dense(a.ctypes.data, a.size, k, table.ctypes.data, C.byref(inv))
    return table, int(inv.value)


def count_dense_range(data, k, win_begin, win_end, table=None):
    a = _bytes_arr(data)
    if table is None:
        table = np.zeros(num_kmers(k), dtype=np.uint32)
    inv = C.c_uint64(0)
    lib().or_count_dense_range(a.ctypes.data, a.size, win_begin, win_end, k, table.ctypes.data, C.byref(inv))
    return table, int(inv.value)


def count_per_seq(data, offsets, k):
    a = _bytes_arr(data)
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offs.size - 1
    sums = np.zeros((num_kmers(k), n), dtype=np.int32)
    inv = np.zeros(max(n, 1), dtype=np.uint64)
    lib().or_count_per_seq(a.ctypes.data, offs.ctypes.data, n, k, sums.ctypes.data, inv.ctypes.data)
    return sums, inv[:n]


def count_sparse(data, k):
    a = _bytes_arr(data)
    kp, cp = C.c_void_p(), C.c_void_p()
    inv = C.c_uint64(0)
    d = lib().or_count_sparse(a.ctypes.data, a.size, k, C.byref(kp), C.byref(cp), C.byref(inv))
    assert d >= 0
    if d == 0:
        keys, counts = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    else:
        keys = np.ctypeslib.as_array(C.cast(kp, C.POINTER(C.c_uint64)), shape=(d,)).copy()
        counts = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_uint32)), shape=(d,)).copy()
    if kp.value:
        lib().or_free(kp)
    if cp.value:
        lib().or_free(cp)
    return keys, counts, int(inv.value)


def triangular_index(i, j, n):
    return int(lib().or_triangular_index(i, j, n))


def distance(sums, offsets, k):
    sums = np.ascontiguousarray(sums, dtype=np.int32)
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offs.size - 1
    out = np.zeros(max(n * (n - 1) // 2, 1), dtype=np.float32)
    lib().or_distance(sums.ctypes.data, offs.ctypes.data, n, k, out.ctypes.data)
    return out[: n * (n - 1) // 2]


def import_seqs_mem(fasta, mode=0, max_seqs=0):
    if isinstance(fasta, str):
        fasta = fasta.encode("latin-1")
    dp, op, ip = C.c_void_p(), C.c_void_p(), C.c_void_p()
    dl, ns, ni = C.c_uint64(), C.c_uint32(), C.c_uint32()
    lib().or_import_seqs_mem(fasta, len(fasta), mode, max_seqs, C.byref(dp), C.byref(dl), C.byref(op),
                             C.byref(ns), C.byref(ip), C.byref(ni))
    data = C.string_at(dp, dl.value)
    offsets = np.ctypeslib.as_array(C.cast(op, C.POINTER(C.c_int64)), shape=(ns.value + 1,)).copy()
    ids_raw = C.string_at(ip)
    ids = ids_raw.decode("latin-1").split("\n")[:-1] if ni.value else []
    for p in (dp, op, ip):
        lib().or_free(p)
    return {"data": data, "offsets": offsets, "ids": ids, "num_seqs": int(ns.value)}


def dump_counts(sums, k, num_seqs):
    sums = np.ascontiguousarray(sums, dtype=np.int32)
    p = lib().or_dump_counts(sums.ctypes.data, k, num_seqs)
    s = C.string_at(p)
    lib().or_free(p)
    return s


def sm64(x):
    return int(lib().or_sm64(x & 0xFFFFFFFFFFFFFFFF))


def mix64(x):
    return int(lib().or_mix64(x & 0xFFFFFFFFFFFFFFFF))


def gen_bases(seed, pos0, n):
    out = np.empty(n, dtype=np.uint8)
    lib().or_gen_bases(seed, pos0, n, out.ctypes.data)
    return out


def gen_genome(seed, total_len, long_runs, short_runs, k, pos0, n):
    out = np.empty(n, dtype=np.uint8)
    lib().or_gen_genome(seed, total_len, long_runs, short_runs, k, pos0, n, out.ctypes.data)
    return out


def gen_reads(seed, genome_len, read_len, err_den, read0, nreads):
    out = np.empty(nreads * (read_len + 1), dtype=np.uint8)
    lib().or_gen_reads(seed, genome_len, read_len, err_den, read0, nreads, out.ctypes.data)
    return out


def fnv1a64(buf):
    a = np.ascontiguousarray(buf)
    return int(lib().or_fnv1a64(a.ctypes.data, a.nbytes))


# ----------------------------------------------------------------------------
# The reference's own host code (oracle/_ref), K = 3..6.  None when not built.
# ----------------------------------------------------------------------------
class Ref:
    def __init__(self, k, L):
        self.k, self.L = k, L
        L.ref_k.restype = C.c_int
        L.ref_count_all.argtypes = [C.c_char_p, C.c_long, C.c_void_p]
        L.ref_permutation.argtypes = [C.c_void_p]
        L.ref_map_size.restype = C.c_long
        L.ref_import.argtypes = [C.c_char_p, C.c_int]
        L.ref_size_all_seqs.restype = C.c_uint
        L.ref_data.restype = C.c_void_p
        L.ref_offset.argtypes = [C.c_int]
        L.ref_id.restype = C.c_char_p
        L.ref_id.argtypes = [C.c_int]
        L.ref_seq.restype = C.c_void_p
        L.ref_seq.argtypes = [C.c_int]
        L.ref_seq_len.restype = C.c_long
        L.ref_seq_len.argtypes = [C.c_int]
        L.ref_distance.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_long), C.c_int, C.c_void_p]
        L.ref_triangular_index.restype = C.c_long
        L.ref_triangular_index.argtypes = [C.c_long, C.c_long, C.c_long]
        assert L.ref_k() == k
        L.ref_init()

    def permutation(self):
        n = 4 ** self.k
        flat = C.create_string_buffer(n * (self.k + 1))
        self.L.ref_permutation(flat)
        raw = flat.raw
        return [raw[i * (self.k + 1): i * (self.k + 1) + self.k].decode() for i in range(n)]

    def count_all(self, seq):
        a = _bytes_arr(seq).tobytes()
        out = np.zeros(4 ** self.k + 1, dtype=np.int32)
        self.L.ref_count_all(a, len(a), out.ctypes.data)
        return out

    def import_seqs(self, path, mode=0):
        n = self.L.ref_import(path.encode(), mode)
        noff = self.L.ref_num_offsets()
        size = self.L.ref_size_all_seqs()
        return {
            "num_seqs": n,
            "ids": [self.L.ref_id(i).decode("latin-1") for i in range(self.L.ref_num_ids())],
            "offsets": np.array([self.L.ref_offset(i) for i in range(noff)], dtype=np.int64),
            "data": C.string_at(self.L.ref_data(), size) if size else b"",
            "seqs": [C.string_at(self.L.ref_seq(i), self.L.ref_seq_len(i)) for i in range(n)],
        }

    def distance(self, seqs):
        bs = [_bytes_arr(s).tobytes() for s in seqs]
        n = len(bs)
        arr = (C.c_char_p * n)(*bs)
        lens = (C.c_long * n)(*[len(b) for b in bs])
        out = np.zeros(max(n * (n - 1) // 2, 1), dtype=np.float32)
        self.L.ref_distance(arr, lens, n, out.ctypes.data)
        return out[: n * (n - 1) // 2]

    def triangular_index(self, i, j, n):
        return int(self.L.ref_triangular_index(i, j, n))


def ref(k):
    """The reference's own code compiled for K=k, or None if oracle/_ref lacks it."""
    if k not in _REF:
        path = os.path.join(HERE, "_ref", "libref_k%d.so" % k)
        _REF[k] = Ref(k, C.CDLL(path)) if os.path.exists(path) else None
    return _REF[k]

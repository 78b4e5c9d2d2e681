"""kmerb200 — thin ctypes harness over libkmerb200.so (include/kmer_b200.h).

Python is plumbing only: device buffers are torch tensors (or raw pointers), the
work is done by the sm_100a kernels behind the C ABI.  There is no CPU fallback:
`Context()` raises if the library or a Blackwell GPU is missing.
"""
import ctypes as C
import os
import re

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO_ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libkmerb200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "kmer_b200.h")

KC_OK = 0
KC_ERR_INVALID, KC_ERR_CUDA, KC_ERR_IO, KC_ERR_NOMEM, KC_ERR_TABLE_FULL, KC_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
DENSE_AUTO, DENSE_DIRECT, DENSE_PARTITION = 0, 1, 2
DENSE_SMEM16C, DENSE_PARTITION_DEFER, DENSE_PARTITION_PAIR, DENSE_PARTITION_TRIO = 3, 4, 5, 6
DENSE_PARTITION_WIDE = 7  # seven windows per record, first-generation scatter (2.72 + 0.82 ms); not chosen by DENSE_AUTO
DENSE_PARTITION_DEFER_PAIR, DENSE_PARTITION_DEFER_TRIO = 8, 9  # the scatter of 4 with the count of 5 / 6
DENSE_PARTITION_WIDE2 = 10  # seven windows per record, second-generation scatter: DENSE_AUTO's choice at k = 12 (1.91 + 0.82 ms)
SPARSE_HASH, SPARSE_SORT, SPARSE_RADIX = 0, 1, 2
SPARSE_AUTO = 3  # the engine picks (radix wherever it exists, by measurement on B200)
SPARSE_UNSORTED = 0x100
SPARSE_NO_FALLBACK = 0x200
IMPORT_BLANKLINE, IMPORT_NONL = 0, 1
MAX_K, MAX_DENSE_K = 31, 16


class KmerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kmerb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def declared_symbols():
    """Every KC_API function the public header declares."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"KC_API\s+[^;(]*?\b(kc_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KmerError(KC_ERR_UNSUPPORTED, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i64, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int
    sig = {
        "kc_version": (i32, []),
        "kc_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "kc_ctx_destroy": (None, [vp]),
        "kc_last_error": (C.c_char_p, [vp]),
        "kc_ctx_device": (i32, [vp]),
        "kc_ctx_sm_count": (i32, [vp]),
        "kc_ctx_launch_count": (u64, [vp]),
        "kc_ctx_last_h2d_bytes": (u64, [vp]),
        "kc_ctx_synchronize": (i32, [vp]),
        "kc_ctx_set_timing": (i32, [vp, i32]),
        "kc_ctx_pass_times": (i32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
        "kc_device_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "kc_device_free": (i32, [vp, vp]),
        "kc_host_alloc_pinned": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "kc_host_free_pinned": (i32, [vp, vp]),
        "kc_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "kc_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "kc_memset_d": (i32, [vp, vp, i32, C.c_size_t]),
        "kc_num_kmers": (u64, [i32]),
        "kc_permutation": (i32, [C.c_char_p, i32, C.POINTER(C.c_char_p)]),
        "kc_kmer_index": (i32, [C.c_char_p, i32, C.POINTER(u64)]),
        "kc_kmer_string": (i32, [u64, i32, C.c_char_p]),
        "kc_import_seqs": (i32, [C.c_char_p, i32, C.c_long, C.POINTER(vp)]),
        "kc_import_seqs_mem": (i32, [C.c_char_p, C.c_size_t, i32, C.c_long, C.POINTER(vp)]),
        "kc_import_seqs_mem_threads": (i32, [C.c_char_p, C.c_size_t, i32, i32, C.POINTER(vp)]),
        "kc_seqset_free": (None, [vp]),
        "kc_seqset_num_seqs": (u32, [vp]),
        "kc_seqset_num_ids": (u32, [vp]),
        "kc_seqset_nbytes": (u64, [vp]),
        "kc_seqset_data": (vp, [vp]),
        "kc_seqset_offsets": (vp, [vp]),
        "kc_seqset_id": (C.c_char_p, [vp, u32]),
        "kc_seqset_to_device": (i32, [vp, vp, C.POINTER(vp), C.POINTER(vp)]),
        "kc_count_per_seq": (i32, [vp, vp, vp, u32, i32, vp]),
        "kc_count_per_seq_async": (i32, [vp, vp, vp, u32, i32, vp, vp]),
        "kc_count_dense": (i32, [vp, vp, u64, i32, vp]),
        "kc_count_dense_async": (i32, [vp, vp, u64, i32, vp, vp]),
        "kc_count_dense_range_async": (i32, [vp, vp, u64, u64, u64, i32, vp, i32, vp]),
        "kc_count_dense_host": (i32, [vp, vp, u64, i32, vp]),
        "kc_count_sparse": (i32, [vp, vp, u64, i32, i32, u64, C.POINTER(vp)]),
        "kc_import_seqs_device": (i32, [vp, vp, C.c_char_p, u64, i32, C.POINTER(vp)]),
        "kc_import_seqs_gpu": (i32, [vp, C.c_char_p, i32, C.POINTER(vp)]),
        "kc_packed_bytes": (u64, [u64]),
        "kc_badmask_bytes": (u64, [u64]),
        "kc_pack_2bit": (i32, [vp, vp, u64, vp, vp, vp]),
        "kc_unpack_2bit": (i32, [vp, vp, vp, u64, vp, vp]),
        "kc_count_dense_packed": (i32, [vp, vp, vp, u64, i32, vp]),
        "kc_pack_2bit_host": (i32, [vp, u64, vp, vp, i32]),
        "kc_pack_2bit_host_body": (i32, [vp, u64, vp, vp, i32, i32]),
        "kc_host_pack_simd": (i32, []),
        "kc_host_pack_threads": (i32, [i32]),
        "kc_count_dense_host_packed": (i32, [vp, vp, u64, i32, vp, i32]),
        "kc_count_dense_host_packed_dev": (i32, [vp, vp, u64, i32, vp, i32]),
        "kc_sparse_radix_plan": (i32, [vp, u64, i32, C.c_uint32, vp]),
        "kc_sparse_radix_scatter": (i32, [vp, vp, u64, vp, vp, vp]),
        "kc_sparse_radix_count": (i32, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]),
        "kc_sparse_free": (None, [vp]),
        "kc_sparse_size": (u64, [vp]),
        "kc_sparse_d_keys": (vp, [vp]),
        "kc_sparse_d_counts": (vp, [vp]),
        "kc_sparse_copy_to_host": (i32, [vp, vp, vp, vp]),
        "kc_sparse_bucket_by_owner": (i32, [vp, vp, vp, u64, u32, vp, vp, vp]),
        "kc_sparse_merge": (i32, [vp, vp, vp, u64, C.POINTER(vp)]),
        "kc_ctx_set_reusable_bytes": (None, [vp, u64]),
        "kc_ctx_release_memory": (None, [vp]),
        "kc_sparse_radix_plan_rounds": (i32, [vp, u64, i32, u32, u32, vp]),
        "kc_sparse_radix_scatter_round": (i32, [vp, vp, u64, vp, u32, vp, vp]),
        "kc_sparse_radix_count_round": (i32, [vp, vp, u32, vp, vp, u32, u32, u32, C.POINTER(vp)]),
        "kc_sparse_radix_count_round_append": (i32, [vp, vp, u32, vp, vp, u32, u32, u32, C.POINTER(vp)]),
        "kc_sparse_concat": (i32, [vp, vp, u32, C.POINTER(vp)]),
        "kc_mix64": (u64, [u64]),
        "kc_window_fingerprint": (i32, [vp, vp, u64, i32, C.POINTER(u64), C.POINTER(u64)]),
        "kc_sparse_fingerprint": (i32, [vp, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "kc_dense_fingerprint": (i32, [vp, vp, i32, C.POINTER(u64), C.POINTER(u64)]),
        "kc_dump_counts": (i32, [C.c_char_p, vp, i32, u32]),
        "kc_kmer_distance": (i32, [vp, vp, vp, u32, i32, vp]),
        "kc_triangular_index": (i64, [i64, i64, i64]),
        "kc_dump_distances": (i32, [C.c_char_p, vp, u64]),
        "kc_gen_bases": (i32, [vp, u64, u64, u64, vp, vp]),
        "kc_gen_genome": (i32, [vp, u64, u64, u32, u32, i32, u64, u64, vp, vp]),
        "kc_gen_reads": (i32, [vp, u64, u64, u32, u32, u64, u64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(x):
    """device/host pointer of a torch tensor, numpy array, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    raise TypeError("cannot take a pointer of %r" % type(x))


def _stream_handle(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


# --------------------------------------------------------------------------
# ctx-less host API
# --------------------------------------------------------------------------
def _check0(rc):
    if rc != KC_OK:
        raise KmerError(rc, lib().kc_last_error(None).decode("utf-8", "replace"))


def num_kmers(k):
    return int(lib().kc_num_kmers(k))


def permutation(alphabet, k):
    """Reference `permutation(alphabet, length, permutations)` (utils.h:21): list of strings."""
    n = len(alphabet) ** k
    bufs = [C.create_string_buffer(k + 1) for _ in range(n)]
    arr = (C.c_char_p * n)(*[C.cast(b, C.c_char_p) for b in bufs])
    _check0(lib().kc_permutation(alphabet.encode(), k, arr))
    return [b.value.decode() for b in bufs]


def kmer_index(s):
    out = C.c_uint64()
    _check0(lib().kc_kmer_index(s.encode("latin-1"), len(s), C.byref(out)))
    return int(out.value)


def kmer_string(idx, k):
    buf = C.create_string_buffer(k + 1)
    _check0(lib().kc_kmer_string(idx, k, buf))
    return buf.value.decode()


def mix64(x):
    return int(lib().kc_mix64(x & 0xFFFFFFFFFFFFFFFF))


def triangular_index(i, j, n):
    return int(lib().kc_triangular_index(i, j, n))


def dump_counts(path, sums, k, num_seqs):
    sums = np.ascontiguousarray(sums, dtype=np.int32)
    _check0(lib().kc_dump_counts(path.encode() if path else None, sums.ctypes.data, k, num_seqs))


def dump_distances(path, dist):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    _check0(lib().kc_dump_distances(path.encode() if path else None, dist.ctypes.data, dist.size))


def pack_2bit_host(h_data, nthreads=0, body=0):
    """host-side packer of the 2-bit store (numpy uint8 in) -> (packed uint8[(n+3)/4], badmask uint32[(n+31)/32]);
    body: 0 = the best loop body of this host, 1 = scalar, 2 = AVX2 (tests compare them)"""
    a = np.ascontiguousarray(h_data, dtype=np.uint8)
    n = a.size
    packed = np.zeros(int(lib().kc_packed_bytes(n)), dtype=np.uint8)
    mask = np.zeros(int(lib().kc_badmask_bytes(n)) // 4, dtype=np.uint32)
    _check0(lib().kc_pack_2bit_host_body(_ptr(a), n, _ptr(packed), _ptr(mask), nthreads, body))
    return packed, mask


def pass_times(ctx):
    """(first_ms, second_ms) of the last dense call when ctx timing is enabled."""
    a, b = C.c_float(), C.c_float()
    ctx._check(lib().kc_ctx_pass_times(ctx._h, C.byref(a), C.byref(b)))
    return float(a.value), float(b.value)


class SeqSet:
    """Result of importSeqs / importSeqsNoNL (main.cu:474 / 401)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_file(cls, path, mode=IMPORT_BLANKLINE, max_seqs=0):
        h = C.c_void_p()
        _check0(lib().kc_import_seqs(path.encode(), mode, max_seqs, C.byref(h)))
        return cls(h)

    @classmethod
    def from_memory(cls, fasta, mode=IMPORT_BLANKLINE, max_seqs=0):
        if isinstance(fasta, str):
            fasta = fasta.encode("latin-1")
        h = C.c_void_p()
        _check0(lib().kc_import_seqs_mem(fasta, len(fasta), mode, max_seqs, C.byref(h)))
        return cls(h)

    @classmethod
    def from_memory_threads(cls, fasta, mode=IMPORT_BLANKLINE, nthreads=0):
        """the multi-threaded parser (no record limit); nthreads <= 0: one per core"""
        if isinstance(fasta, str):
            fasta = fasta.encode("latin-1")
        h = C.c_void_p()
        _check0(lib().kc_import_seqs_mem_threads(fasta, len(fasta), mode, nthreads, C.byref(h)))
        return cls(h)

    @classmethod
    def from_device(cls, ctx, d_raw, h_raw, nbytes, mode=IMPORT_BLANKLINE):
        """parse a FASTA file image that already sits in device memory, on the GPU (f2); h_raw = the same
        bytes on the host (for the id strings) or None"""
        h = C.c_void_p()
        ctx._check(lib().kc_import_seqs_device(ctx._h, _ptr(d_raw), h_raw, nbytes, mode, C.byref(h)))
        return cls(h)

    @classmethod
    def from_file_gpu(cls, ctx, path, mode=IMPORT_BLANKLINE):
        """the file goes to the device as it is and is parsed there (f2)"""
        h = C.c_void_p()
        ctx._check(lib().kc_import_seqs_gpu(ctx._h, path.encode(), mode, C.byref(h)))
        return cls(h)

    @property
    def num_seqs(self):
        return int(lib().kc_seqset_num_seqs(self._h))

    @property
    def ids(self):
        return [lib().kc_seqset_id(self._h, i).decode("latin-1") for i in range(lib().kc_seqset_num_ids(self._h))]

    @property
    def nbytes(self):
        return int(lib().kc_seqset_nbytes(self._h))

    @property
    def data(self):
        n = self.nbytes
        return C.string_at(lib().kc_seqset_data(self._h), n) if n else b""

    @property
    def offsets(self):
        p = C.cast(lib().kc_seqset_offsets(self._h), C.POINTER(C.c_int64))
        return np.ctypeslib.as_array(p, shape=(self.num_seqs + 1,)).copy()

    def to_device(self, ctx):
        d, o = C.c_void_p(), C.c_void_p()
        ctx._check(lib().kc_seqset_to_device(ctx._h, self._h, C.byref(d), C.byref(o)))
        return d.value, o.value

    def close(self):
        if self._h:
            lib().kc_seqset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RadixPlan(C.Structure):
    """kc_radix_plan (include/kmer_b200.h)"""
    _fields_ = [("k", C.c_int32), ("world", C.c_uint32), ("partitions", C.c_uint32), ("parts_per_rank", C.c_uint32),
                ("grid", C.c_uint32), ("rec_bytes", C.c_uint32), ("shape", C.c_uint32), ("round_bits", C.c_uint32),
                ("max_windows", C.c_uint64), ("region_records", C.c_uint64), ("slab_bytes", C.c_uint64),
                ("counts_bytes", C.c_uint64)]


class Sparse:
    def __init__(self, ctx, handle):
        self._ctx, self._h = ctx, handle

    def __len__(self):
        return int(lib().kc_sparse_size(self._h))

    @property
    def d_keys(self):
        return lib().kc_sparse_d_keys(self._h)

    @property
    def d_counts(self):
        return lib().kc_sparse_d_counts(self._h)

    def to_host(self):
        n = len(self)
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        self._ctx._check(lib().kc_sparse_copy_to_host(self._ctx._h, self._h, keys.ctypes.data, counts.ctypes.data))
        return keys, counts

    def close(self):
        if self._h:
            lib().kc_sparse_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU.  Raises KmerError when no sm_100 device / library is present."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().kc_ctx_create(device, C.byref(h))
        if rc != KC_OK:
            raise KmerError(rc, lib().kc_last_error(None).decode("utf-8", "replace"))
        self._h = h
        self.device = device

    def _check(self, rc):
        if rc != KC_OK:
            raise KmerError(rc, lib().kc_last_error(self._h).decode("utf-8", "replace"))

    def close(self):
        if self._h:
            lib().kc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return lib().kc_ctx_sm_count(self._h)

    @property
    def launch_count(self):
        return int(lib().kc_ctx_launch_count(self._h))

    @property
    def last_h2d_bytes(self):
        """bytes the last count_dense_host[_packed] call copied host -> device"""
        return int(lib().kc_ctx_last_h2d_bytes(self._h))

    def synchronize(self):
        self._check(lib().kc_ctx_synchronize(self._h))

    # ---- torch helpers -------------------------------------------------
    def _torch(self):
        import torch
        return torch

    def empty(self, shape, dtype):
        torch = self._torch()
        return torch.empty(shape, dtype=dtype, device="cuda:%d" % self.device)

    # ---- counting ---------------------------------------------------------
    def count_dense(self, d_data, nbytes, k, table=None, stream=None, sync=True):
        """uint32[4^k] table (torch int32 tensor viewed bit-exactly) of a device byte buffer."""
        torch = self._torch()
        if table is None:
            table = torch.empty(num_kmers(k), dtype=torch.int32, device="cuda:%d" % self.device)
        self._check(lib().kc_count_dense_async(self._h, _ptr(d_data), nbytes, k, _ptr(table), _stream_handle(stream)))
        if sync:
            torch.cuda.current_stream().synchronize() if stream is None else stream.synchronize()
        return table

    def count_dense_range(self, d_data, nbytes, win_begin, win_end, k, table, algo=DENSE_AUTO, stream=None):
        self._check(lib().kc_count_dense_range_async(self._h, _ptr(d_data), nbytes, win_begin, win_end, k, _ptr(table),
                                                     algo, _stream_handle(stream)))
        return table

    def count_dense_host(self, h_data, k, h_table=None):
        """End-to-end from a host buffer (numpy uint8 / pinned torch tensor)."""
        n = h_data.numel() if hasattr(h_data, "numel") else h_data.size
        if h_table is None:
            h_table = np.empty(num_kmers(k), dtype=np.uint32)
        self._check(lib().kc_count_dense_host(self._h, _ptr(h_data), n, k, _ptr(h_table)))
        return h_table

    def count_dense_host_packed(self, h_data, k, h_table=None, nthreads=0):
        """count_dense_host through the packed form: host threads pack (0.375 B/base over PCIe), the GPU unpacks and counts"""
        n = h_data.numel() if hasattr(h_data, "numel") else h_data.size
        if h_table is None:
            h_table = np.empty(num_kmers(k), dtype=np.uint32)
        self._check(lib().kc_count_dense_host_packed(self._h, _ptr(h_data), n, k, _ptr(h_table), nthreads))
        return h_table

    def count_dense_host_packed_dev(self, h_data, k, d_table, nthreads=0):
        """count_dense_host_packed into a device table (overwritten), synchronous"""
        n = h_data.numel() if hasattr(h_data, "numel") else h_data.size
        self._check(lib().kc_count_dense_host_packed_dev(self._h, _ptr(h_data), n, k, _ptr(d_table), nthreads))
        return d_table

    def count_per_seq(self, d_data, d_offsets, num_seqs, k, sums=None, stream=None, sync=True):
        torch = self._torch()
        if sums is None:
            sums = torch.empty((num_kmers(k), num_seqs), dtype=torch.int32, device="cuda:%d" % self.device)
        self._check(lib().kc_count_per_seq_async(self._h, _ptr(d_data), _ptr(d_offsets), num_seqs, k, _ptr(sums),
                                                 _stream_handle(stream)))
        if sync:
            torch.cuda.current_stream().synchronize() if stream is None else stream.synchronize()
        return sums

    def count_sparse(self, d_data, nbytes, k, algo=SPARSE_HASH, capacity_hint=0):
        self._torch().cuda.current_stream().synchronize()
        h = C.c_void_p()
        self._check(lib().kc_count_sparse(self._h, _ptr(d_data), nbytes, k, algo, capacity_hint, C.byref(h)))
        return Sparse(self, h)

    # ---- 2-bit packed store ("next" row f4) ------------------------------------------------
    def pack_2bit(self, d_data, nbytes):
        """-> (packed uint8[(n+3)/4], badmask int32[(n+31)/32]) on this GPU"""
        torch = self._torch()
        dev = "cuda:%d" % self.device
        packed = torch.empty(int(lib().kc_packed_bytes(nbytes)) + 4, dtype=torch.uint8, device=dev)
        mask = torch.empty(int(lib().kc_badmask_bytes(nbytes)) // 4 + 1, dtype=torch.int32, device=dev)
        self._check(lib().kc_pack_2bit(self._h, _ptr(d_data), nbytes, _ptr(packed), _ptr(mask), _stream_handle(None)))
        torch.cuda.current_stream().synchronize()
        return packed, mask

    def unpack_2bit(self, packed, mask, nbases):
        torch = self._torch()
        out = torch.empty(nbases, dtype=torch.uint8, device="cuda:%d" % self.device)
        self._check(lib().kc_unpack_2bit(self._h, _ptr(packed), _ptr(mask), nbases, _ptr(out), _stream_handle(None)))
        torch.cuda.current_stream().synchronize()
        return out

    def count_dense_packed(self, packed, mask, nbases, k, table=None):
        torch = self._torch()
        torch.cuda.current_stream().synchronize()
        if table is None:
            table = torch.empty(num_kmers(k), dtype=torch.int32, device="cuda:%d" % self.device)
        self._check(lib().kc_count_dense_packed(self._h, _ptr(packed), _ptr(mask), nbases, k, _ptr(table)))
        return table

    # ---- stages of KC_SPARSE_RADIX (the multi-GPU path runs an all-to-all between them) ----
    def release_memory(self):
        """give back the ctx's scratch areas and the idle blocks of the device memory pool"""
        lib().kc_ctx_release_memory(self._h)

    def radix_plan(self, max_windows, k, world, min_round_bits=0):
        plan = RadixPlan()
        try:  # what torch's caching allocator holds free counts as available for the slab tensors
            torch = self._torch()
            lib().kc_ctx_set_reusable_bytes(self._h, max(0, torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)))
        except Exception:
            pass
        self._check(lib().kc_sparse_radix_plan_rounds(self._h, max_windows, k, world, min_round_bits, C.addressof(plan)))
        return plan

    def radix_scatter(self, d_data, nbytes, plan, rnd=0, out=None):
        """-> (slabs uint8[plan.slab_bytes], counts int32[plan.counts_bytes / 4]) on this GPU, partition-major, for
        round `rnd` of the plan (`out` = buffers of an earlier round to reuse); raises KmerError(KC_ERR_TABLE_FULL)
        when a region overflowed"""
        torch = self._torch()
        torch.cuda.current_stream().synchronize()
        dev = "cuda:%d" % self.device
        if out is None:
            slabs = torch.empty(plan.slab_bytes, dtype=torch.uint8, device=dev)
            counts = torch.empty(plan.counts_bytes // 4, dtype=torch.int32, device=dev)
        else:
            slabs, counts = out
        torch.cuda.current_stream().synchronize()
        self._check(lib().kc_sparse_radix_scatter_round(self._h, _ptr(d_data), nbytes, C.addressof(plan), rnd, _ptr(slabs), _ptr(counts)))
        return slabs, counts

    def radix_count(self, plan, slabs, counts, nsrc, part_first, nparts, rnd=0):
        self._torch().cuda.current_stream().synchronize()
        h = C.c_void_p()
        self._check(lib().kc_sparse_radix_count_round(self._h, C.addressof(plan), rnd, _ptr(slabs), _ptr(counts), nsrc, part_first,
                                                      nparts, C.byref(h)))
        return Sparse(self, h)

    def radix_count_append(self, plan, slabs, counts, nsrc, part_first, nparts, rnd, acc=None):
        """count round `rnd` (rounds in ascending order) behind what `acc` holds; returns the one growing result"""
        self._torch().cuda.current_stream().synchronize()
        h = C.c_void_p(acc._h.value if acc is not None else None)
        try:
            self._check(lib().kc_sparse_radix_count_round_append(self._h, C.addressof(plan), rnd, _ptr(slabs), _ptr(counts), nsrc,
                                                                 part_first, nparts, C.byref(h)))
        finally:
            if acc is not None:
                acc._h = h
        return acc if acc is not None else Sparse(self, h)

    def sparse_concat(self, parts):
        """one result from ascending pieces (the rounds of a radix plan); the pieces are closed"""
        if len(parts) == 1:
            return parts[0]
        arr = (C.c_void_p * len(parts))(*[p._h for p in parts])
        h = C.c_void_p()
        self._check(lib().kc_sparse_concat(self._h, arr, len(parts), C.byref(h)))
        for p in parts:
            p.close()
        return Sparse(self, h)

    # ---- full-scale self-checks (csrc/check.cu): (fingerprint, total) pairs; a correct count has equal pairs ----
    def window_fingerprint(self, d_data, nbytes, k):
        """(sum of mix64(code) over the valid windows of the input mod 2^64, number of valid windows)"""
        fp, n = C.c_uint64(), C.c_uint64()
        self._check(lib().kc_window_fingerprint(self._h, _ptr(d_data), nbytes, k, C.byref(fp), C.byref(n)))
        return int(fp.value), int(n.value)

    def sparse_fingerprint(self, sp):
        """(sum of count * mix64(code) over a sparse result mod 2^64, sum of counts, positions where the keys are
        not strictly ascending — 0 for a valid result)"""
        fp, n, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(lib().kc_sparse_fingerprint(self._h, sp._h, C.byref(fp), C.byref(n), C.byref(d)))
        return int(fp.value), int(n.value), int(d.value)

    def dense_fingerprint(self, table, k):
        """(sum of count * mix64(code) over a dense table mod 2^64, sum of counts)"""
        fp, n = C.c_uint64(), C.c_uint64()
        self._check(lib().kc_dense_fingerprint(self._h, _ptr(table), k, C.byref(fp), C.byref(n)))
        return int(fp.value), int(n.value)

    def sparse_merge(self, d_keys, d_counts, n):
        self._torch().cuda.current_stream().synchronize()
        h = C.c_void_p()
        self._check(lib().kc_sparse_merge(self._h, _ptr(d_keys), _ptr(d_counts), n, C.byref(h)))
        return Sparse(self, h)

    def sparse_bucket_by_owner(self, d_keys, d_counts, n, num_owners, d_keys_out, d_counts_out):
        self._torch().cuda.current_stream().synchronize()
        sizes = np.zeros(num_owners, dtype=np.uint64)
        self._check(lib().kc_sparse_bucket_by_owner(self._h, _ptr(d_keys), _ptr(d_counts), n, num_owners,
                                                    _ptr(d_keys_out), _ptr(d_counts_out), sizes.ctypes.data))
        return sizes

    def kmer_distance(self, d_sums, d_offsets, num_seqs, k, dist=None):
        torch = self._torch()
        torch.cuda.current_stream().synchronize()
        npairs = num_seqs * (num_seqs - 1) // 2
        if dist is None:
            dist = torch.zeros(max(npairs, 1), dtype=torch.float32, device="cuda:%d" % self.device)
        self._check(lib().kc_kmer_distance(self._h, _ptr(d_sums), _ptr(d_offsets), num_seqs, k, _ptr(dist)))
        return dist[:npairs]

    # ---- synthetic inputs ---------------------------------------------------
    def gen_bases(self, seed, pos0, n, out=None, stream=None):
        torch = self._torch()
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % self.device)
        self._check(lib().kc_gen_bases(self._h, seed, pos0, n, _ptr(out), _stream_handle(stream)))
        return out

    def gen_genome(self, seed, total_len, long_runs, short_runs, k, pos0, n, out=None, stream=None):
        torch = self._torch()
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device="cuda:%d" % self.device)
        self._check(lib().kc_gen_genome(self._h, seed, total_len, long_runs, short_runs, k, pos0, n, _ptr(out),
                                        _stream_handle(stream)))
        return out

    def gen_reads(self, seed, genome_len, read_len, err_den, read0, nreads, out=None, stream=None):
        torch = self._torch()
        if out is None:
            out = torch.empty(nreads * (read_len + 1), dtype=torch.uint8, device="cuda:%d" % self.device)
        self._check(lib().kc_gen_reads(self._h, seed, genome_len, read_len, err_den, read0, nreads, _ptr(out),
                                       _stream_handle(stream)))
        return out


# --------------------------------------------------------------------------
# multi-GPU sharding logic (pure host arithmetic; tested on CPU with gloo)
# --------------------------------------------------------------------------
def shard_windows(nbytes, k, rank, world):
    """Window-start range [b, e) of `rank` and the byte range it must be able to
    read: its own bytes plus a (k-1)-byte halo (dense mode)."""
    nwin = max(nbytes - k + 1, 0)
    b = nwin * rank // world
    e = nwin * (rank + 1) // world
    return b, e, b, min(e + k - 1, nbytes) if e > b else b


def shard_seqs(offsets, rank, world):
    """Sequence range [s0, s1) of `rank` for the per-sequence table (independent sequences): cut at
    sequence starts, as close as they come to equal BYTES per rank (offsets = the loader's num_seqs + 1 starts)."""
    off = np.asarray(offsets, dtype=np.int64)
    n = off.size - 1
    if n <= 0:
        return 0, 0
    total = int(off[-1] - off[0])
    cut = [int(np.searchsorted(off[:-1], off[0] + total * r // world, side="left")) for r in range(world + 1)]
    cut[0], cut[world] = 0, n
    for r in range(1, world + 1):   # monotone
        cut[r] = max(cut[r], cut[r - 1])
    return cut[rank], cut[rank + 1]


def shard_reads(nreads, rank, world):
    """Read range of `rank` for the sparse path (reads never straddle shards)."""
    return nreads * rank // world, nreads * (rank + 1) // world

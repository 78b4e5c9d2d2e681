// kmer_b200.hpp — C++ host-side mirror of the reference's entry points for the
// counting path, layered on the C ABI (include/kmer_b200.h).  Same names and
// argument meaning as the reference so a maintainer can swap call sites 1:1:
//
//   reference (one TU, globals)                      here
//   ------------------------------------------------ -----------------------------------
//   permutation(alphabet, length, perms) utils.h:21  kmerb200::permutation(...)
//   importSeqs(file)      main.cu:474 -> globals     kmerb200::importSeqs(file)      -> Sequences
//   importSeqsNoNL(file)  main.cu:401                kmerb200::importSeqsNoNL(file)  -> Sequences
//   sumKmereCoincidencesGlobalMemory<<<..>>>(data, indices, num_seqs, sums)
//                         kernels.h:113, main.cu:290 Engine::sumKmereCoincidences(seqs, k) -> sums
//   minKmeres2<<<..>>> x num_seqs  main.cu:327-335   Engine::minKmeres(sums, seqs, k)     -> mins
//   getIdxTriangularMatrixRowMajor kernels.h:46      kmerb200::getIdxTriangularMatrixRowMajor
//
// Error behaviour: the reference printf()s and exit()s (main.cu:224-227,477-480);
// here every failure throws kmerb200::Error carrying the kc_status and message.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kmer_b200.h"

namespace kmerb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc, const kc_ctx* ctx = nullptr) {
    if (rc != KC_OK) throw Error(rc, kc_last_error(ctx));
}

inline void permutation(const char* alphabet, int length, char** permutations) {
    check(kc_permutation(alphabet, length, permutations));
}

inline long getIdxTriangularMatrixRowMajor(long i, long j, long n) { return (long)kc_triangular_index(i, j, n); }

// what importSeqs leaves in the reference's globals (main.cu:65-70,34-35)
struct Sequences {
    kc_seqset* handle = nullptr;
    std::vector<std::string> ids;
    int numberOfSequenses = 0;        // sic, main.cu:34
    uint64_t size_all_seqs = 0;       // main.cu:35
    const char* data = nullptr;       // host image of the reference's managed `data`
    const int64_t* indexes = nullptr; // numberOfSequenses + 1 offsets
    Sequences() = default;
    Sequences(const Sequences&) = delete;
    Sequences& operator=(const Sequences&) = delete;
    Sequences(Sequences&& o) noexcept { *this = std::move(o); }
    Sequences& operator=(Sequences&& o) noexcept {
        std::swap(handle, o.handle);
        ids.swap(o.ids);
        numberOfSequenses = o.numberOfSequenses;
        size_all_seqs = o.size_all_seqs;
        data = o.data;
        indexes = o.indexes;
        return *this;
    }
    ~Sequences() { kc_seqset_free(handle); }
    long length(int i) const { return (long)(indexes[i + 1] - indexes[i] - 1); }
};

inline Sequences import_impl(const std::string& file, int mode, long max_seqs) {
    Sequences s;
    check(kc_import_seqs(file.c_str(), mode, max_seqs, &s.handle));
    s.numberOfSequenses = (int)kc_seqset_num_seqs(s.handle);
    s.size_all_seqs = kc_seqset_nbytes(s.handle);
    s.data = kc_seqset_data(s.handle);
    s.indexes = kc_seqset_offsets(s.handle);
    for (uint32_t i = 0; i < kc_seqset_num_ids(s.handle); i++) s.ids.emplace_back(kc_seqset_id(s.handle, i));
    return s;
}
// MAX_SEQS is 100 in the reference (main.cu:30); pass 0 for "no limit"
inline Sequences importSeqs(const std::string& inputFile, long max_seqs = 100) { return import_impl(inputFile, KC_IMPORT_BLANKLINE, max_seqs); }
inline Sequences importSeqsNoNL(const std::string& inputFile, long max_seqs = 100) { return import_impl(inputFile, KC_IMPORT_NONL, max_seqs); }

class Engine;
// f2, device side: the file goes to the GPU as it is and is parsed there (kc_import_seqs_gpu); the
// returned Sequences has no host image of `data` unless somebody asks kc_seqset_data for it
inline Sequences importSeqsGpu(Engine& eng, const std::string& inputFile, bool nonl = false);

class Engine {
  public:
    explicit Engine(int device = 0) { check(kc_ctx_create(device, &ctx_)); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    ~Engine() { kc_ctx_destroy(ctx_); }
    kc_ctx* ctx() const { return ctx_; }

    // per-sequence count table, kmer-major: sums[entry + num_seqs * kmer] (kernels.h:142)
    std::vector<int32_t> sumKmereCoincidences(Sequences& seqs, int k, int32_t** d_sums_out = nullptr) {
        const char* d_data;
        const int64_t* d_off;
        check(kc_seqset_to_device(ctx_, seqs.handle, &d_data, &d_off), ctx_);
        const size_t n = (size_t)kc_num_kmers(k) * seqs.numberOfSequenses;
        void* d_sums = nullptr;
        check(kc_device_alloc(ctx_, n * sizeof(int32_t), &d_sums), ctx_);
        std::vector<int32_t> sums(n);
        int rc = kc_count_per_seq(ctx_, d_data, d_off, (uint32_t)seqs.numberOfSequenses, k, (int32_t*)d_sums);
        if (rc == KC_OK) rc = kc_memcpy_d2h(ctx_, sums.data(), d_sums, n * sizeof(int32_t));
        if (rc != KC_OK || !d_sums_out) kc_device_free(ctx_, d_sums);
        check(rc, ctx_);
        if (d_sums_out) *d_sums_out = (int32_t*)d_sums;
        return sums;
    }

    // aggregate dense table of a host byte buffer (any non-ACGT byte splits windows)
    // packer_threads < 0: plain copy (1 byte per base over PCIe); >= 0: the host cores pack first
    // (kc_count_dense_host_packed, <= 0.375 bytes per base over PCIe; 0 = as many threads as the process may use)
    std::vector<uint32_t> countDense(const char* h_data, uint64_t nbytes, int k, int packer_threads = -1) {
        std::vector<uint32_t> table((size_t)kc_num_kmers(k));
        if (packer_threads < 0)
            check(kc_count_dense_host(ctx_, h_data, nbytes, k, table.data()), ctx_);
        else
            check(kc_count_dense_host_packed(ctx_, h_data, nbytes, k, table.data(), packer_threads), ctx_);
        return table;
    }

    // sparse count (k <= 31) of a host byte buffer: distinct LE codes ascending + counts.
    // algo: KC_SPARSE_HASH / KC_SPARSE_SORT / KC_SPARSE_RADIX (the last falls back to the hash
    // path by itself when skewed data overflows one of its regions)
    void countSparse(const char* h_data, uint64_t nbytes, int k, int algo, std::vector<uint64_t>& keys,
                     std::vector<uint32_t>& counts) {
        void* d_data = nullptr;
        check(kc_device_alloc(ctx_, nbytes, &d_data), ctx_);
        kc_sparse* sp = nullptr;
        int rc = kc_memcpy_h2d(ctx_, d_data, h_data, nbytes);
        if (rc == KC_OK) rc = kc_count_sparse(ctx_, (const char*)d_data, nbytes, k, algo, 0, &sp);
        if (rc == KC_OK) {
            keys.resize((size_t)kc_sparse_size(sp));
            counts.resize(keys.size());
            rc = kc_sparse_copy_to_host(ctx_, sp, keys.data(), counts.data());
        }
        kc_sparse_free(sp);
        kc_device_free(ctx_, d_data);
        check(rc, ctx_);
    }

    // packed strict upper triangle of k-mer distances (main.cu:327-358)
    std::vector<float> minKmeres(const int32_t* d_sums, Sequences& seqs, int k) {
        const char* d_data;
        const int64_t* d_off;
        check(kc_seqset_to_device(ctx_, seqs.handle, &d_data, &d_off), ctx_);
        const size_t n = seqs.numberOfSequenses;
        const size_t pairs = n * (n + 1) / 2 - n;  // resultsArraySize, main.cu:165
        std::vector<float> mins(pairs);
        if (!pairs) return mins;
        void* d_mins = nullptr;
        check(kc_device_alloc(ctx_, pairs * sizeof(float), &d_mins), ctx_);
        int rc = kc_kmer_distance(ctx_, d_sums, d_off, (uint32_t)n, k, (float*)d_mins);
        if (rc == KC_OK) rc = kc_memcpy_d2h(ctx_, mins.data(), d_mins, pairs * sizeof(float));
        kc_device_free(ctx_, d_mins);
        check(rc, ctx_);
        return mins;
    }

  private:
    kc_ctx* ctx_ = nullptr;
};

inline Sequences importSeqsGpu(Engine& eng, const std::string& inputFile, bool nonl) {
    Sequences s;
    check(kc_import_seqs_gpu(eng.ctx(), inputFile.c_str(), nonl ? KC_IMPORT_NONL : KC_IMPORT_BLANKLINE, &s.handle), eng.ctx());
    s.numberOfSequenses = (int)kc_seqset_num_seqs(s.handle);
    s.size_all_seqs = kc_seqset_nbytes(s.handle);
    s.data = nullptr;  // stays on the device
    s.indexes = kc_seqset_offsets(s.handle);
    for (uint32_t i = 0; i < kc_seqset_num_ids(s.handle); i++) s.ids.emplace_back(kc_seqset_id(s.handle, i));
    return s;
}

}  // namespace kmerb200

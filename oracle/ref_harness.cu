/*
 * ref_harness.cu — drives the REFERENCE's own host functions without a GPU.
 * *** TEST INFRASTRUCTURE ONLY *** (see oracle/kmer_oracle.c header).
 *
 * The reference translation unit (/root/reference/main.cu, with kernels.h and
 * utils.h) is #include'd UNMODIFIED from where it lies — no reference source is
 * copied into this repo.  `main` is renamed so the TU can live in a shared
 * object, and cudaMallocManaged (the only CUDA call on the loader path,
 * main.cu:460,532) is redirected to malloc so importSeqs runs on a CPU-only box.
 * Built once per K (the reference's k is a compile-time macro, kernels.h:11-15)
 * by oracle/build_ref.sh into oracle/_ref/libref_k<K>.so; K = 3..6 are the
 * values the reference compiles for.
 *
 * What is exercised is the reference's code, byte for byte:
 *   permutation()            utils.h:21-50
 *   permutationsMap fill     main.cu:134-135 (restated in ref_init: it is inline in main())
 *   permutationsCountAll()   main.cu:636-646
 *   importSeqs / NoNL        main.cu:474-545 / 401-473
 *   sequentialKmerCount2()   main.cu:587-621
 */
#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
#include <cuda.h>
#include <cooperative_groups.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static cudaError_t harness_managed_alloc(void** p, size_t nbytes) {
    *p = malloc(nbytes ? nbytes : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
#define cudaMallocManaged(p, nbytes) harness_managed_alloc((void**)(p), (nbytes))
#define main reference_main
#include "/root/reference/main.cu"
#undef main
#undef cudaMallocManaged
#undef N /* main.cu:25 defines a macro called N */

static bool g_ready = false;

extern "C" {

__attribute__((visibility("default"))) int ref_k(void) { return K; }

/* what main() does at main.cu:122-135 (buffers calloc'd: the reference never
 * writes the terminator, main.cu:128 / utils.h:31-33) */
__attribute__((visibility("default"))) void ref_init(void) {
    if (g_ready) return;
    const char* alphabet = "ACGT";
    int permsSize = (int)pow(4, K);
    char** p = (char**)malloc(permsSize * sizeof(char*));
    for (int i = 0; i < permsSize; i++) p[i] = (char*)calloc(K + 1, 1);
    permutation(alphabet, K, p);
    for (int i = 0; i < permsSize; i++) permutationsList.push_back(p[i]);
    for (int i = 0; i < PERMS_KMERES; i++) permutationsMap[p[i]] = i + 1;
    g_ready = true;
}

/* flat[i*(K+1) .. ] = i-th enumerated k-mer */
__attribute__((visibility("default"))) void ref_permutation(char* flat) {
    ref_init();
    for (int i = 0; i < PERMS_KMERES; i++) {
        memcpy(flat + (size_t)i * (K + 1), permutationsList[i].c_str(), K);
        flat[(size_t)i * (K + 1) + K] = 0;
    }
}

/* permutationsCountAll on seq + '|' (the '|' is what importSeqs appends,
 * main.cu:505); counts has 4^K + 1 ints, [0] = invalid bucket */
__attribute__((visibility("default"))) void ref_count_all(const char* seq, long len, int* counts) {
    ref_init();
    std::string s(seq, (size_t)len);
    s += "|";
    permutationsCountAll(s, counts, PERMS_KMERES, K);
}

/* size of permutationsMap (it grows on invalid lookups, main.cu:644) */
__attribute__((visibility("default"))) long ref_map_size(void) { return (long)permutationsMap.size(); }

/* run importSeqs (mode 0) or importSeqsNoNL (mode 1) on a file; returns the
 * number of sequences and leaves the results in the reference's globals */
__attribute__((visibility("default"))) int ref_import(const char* path, int mode) {
    ids.clear();
    seqs.clear();
    indexes_aux.clear();
    if (::data) free(::data);
    ::data = nullptr;
    numberOfSequenses = 0;
    size_all_seqs = 0;
    if (mode == 0)
        importSeqs(path);
    else
        importSeqsNoNL(path);
    return numberOfSequenses;
}
__attribute__((visibility("default"))) int ref_num_ids(void) { return (int)ids.size(); }
__attribute__((visibility("default"))) int ref_num_offsets(void) { return (int)indexes_aux.size(); }
__attribute__((visibility("default"))) unsigned ref_size_all_seqs(void) { return size_all_seqs; }
__attribute__((visibility("default"))) const char* ref_data(void) { return ::data; }
__attribute__((visibility("default"))) int ref_offset(int i) { return indexes_aux[i]; }
__attribute__((visibility("default"))) const char* ref_id(int i) { return ids[i].c_str(); }
__attribute__((visibility("default"))) const char* ref_seq(int i) { return seqs[i].c_str(); }
__attribute__((visibility("default"))) long ref_seq_len(int i) { return (long)seqs[i].size(); }

/* sequentialKmerCount2 over caller-supplied sequences (each gets the '|' the
 * loader would append); out has n(n-1)/2 floats in the packed order of
 * main.cu:671-673 */
__attribute__((visibility("default"))) void ref_distance(const char** in, const long* lens, int n,
                                                         float* out) {
    ref_init();
    seqs.clear();
    for (int i = 0; i < n; i++) {
        std::string s(in[i], (size_t)lens[i]);
        s += "|";
        seqs.push_back(s);
    }
    numberOfSequenses = n;
    resultsArraySize = (long)n * (n + 1) / 2 - n;
    distancesSequential = (float*)calloc(resultsArraySize ? resultsArraySize : 1, sizeof(float));
    sequentialKmerCount2(seqs, permutationsList, K);
    for (long i = 0; i < resultsArraySize; i++) out[i] = distancesSequential[i];
    free(distancesSequential);
    distancesSequential = nullptr;
}

__attribute__((visibility("default"))) long ref_triangular_index(long i, long j, long n) {
    return getIdxTriangularMatrixRowMajorSeq(i, j, n);
}
}

/*
 * kmer_b200.h — C ABI of the B200-native k-mer counting engine (libkmerb200.so).
 *
 * This is the drop-in boundary for the k-mer COUNTING path of
 * axlwild/dna-kmeres-parallel.  The reference has no FFI of its own (it is a
 * monolithic main()); each entry point below replaces one host entry point or
 * data contract that the reference's main() uses around the count step, and
 * cites it as  <file>:<line>  into the reference tree.
 *
 * Conventions
 *   - plain C, pointers + sizes, no torch / C++ types in any signature;
 *   - every call returns KC_OK (0) or a negative kc_status; the message is
 *     available from kc_last_error(ctx) (ctx may be NULL for ctx-less calls);
 *   - library code never calls exit() (the reference does: main.cu:224-227);
 *   - "d_" arguments are device pointers on the ctx's GPU, "h_" are host;
 *   - plain calls are synchronous (the reference synchronises after every
 *     launch, main.cu:291); *_async variants enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream) and
 *     return immediately;
 *   - one ctx is not thread-safe; distinct ctxs are.
 *   - the *_async entry points take the caller's stream but share the ctx's scratch (partition slabs, per-CTA
 *     partial tables): calls on ONE ctx must be stream-ordered, one in flight at a time — two async calls on
 *     different streams of one ctx would overwrite each other's slabs.  Scratch grows on demand (free + malloc):
 *     a CUDA graph captured around a call holds the scratch pointer of that moment, so capture after a warm-up
 *     call of the largest size the graph will see (bench.py does), or use one ctx per graph.
 *   - there is NO CPU fallback: every counting call runs sm_100a kernels.
 *
 * K-mer semantics (bit-exact with the reference):
 *   window = k consecutive bytes of ONE sequence; counted iff all k bytes are
 *   upper-case A/C/G/T (main.cu:641-644, kernels.h:133-140); forward strand
 *   only; table index is little-endian in the string,
 *   idx = sum_p code(s[p]) * 4^p, A=0 C=1 G=2 T=3 (utils.h:30-47).
 */
#ifndef KMER_B200_H
#define KMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define KC_API __attribute__((visibility("default")))
#else
#define KC_API
#endif

typedef enum kc_status {
    KC_OK = 0,
    KC_ERR_INVALID = -1,     /* bad argument (k out of range, NULL pointer, ...) */
    KC_ERR_CUDA = -2,        /* CUDA runtime error, text in kc_last_error */
    KC_ERR_IO = -3,          /* file could not be opened / written */
    KC_ERR_NOMEM = -4,       /* host or device allocation failed */
    KC_ERR_TABLE_FULL = -5,  /* sparse table capacity exhausted (caller may retry larger) */
    KC_ERR_UNSUPPORTED = -6  /* not an sm_100 device, dense k too large, ... */
} kc_status;

typedef struct kc_ctx kc_ctx;         /* device, streams, scratch */
typedef struct kc_seqset kc_seqset;   /* loaded FASTA records */
typedef struct kc_sparse kc_sparse;   /* sorted (code,count) result of a sparse count */

/* ------------------------------------------------------------------ */
/* Context + errors.  Replaces the reference's implicit "device 0,     */
/* default stream, printf+exit" conventions (main.cu:150-158,224-227). */
/* ------------------------------------------------------------------ */
KC_API int kc_version(void);
KC_API int kc_ctx_create(int device, kc_ctx** out);
KC_API void kc_ctx_destroy(kc_ctx* ctx);
KC_API const char* kc_last_error(const kc_ctx* ctx);
KC_API int kc_ctx_device(const kc_ctx* ctx);
KC_API int kc_ctx_sm_count(const kc_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench "gpu_launches") */
KC_API uint64_t kc_ctx_launch_count(const kc_ctx* ctx);
/* bytes the last kc_count_dense_host[_packed] call of this ctx copied host -> device */
KC_API uint64_t kc_ctx_last_h2d_bytes(const kc_ctx* ctx);
/* block until everything enqueued through this ctx has finished */
KC_API int kc_ctx_synchronize(kc_ctx* ctx);
/* Measurement aid: when enabled, the dense entry points bracket their kernels
 * with CUDA events on the launching stream; after the stream is synchronised
 * kc_ctx_pass_times returns the device time of the last call's kernels in ms
 * (direct path: *first = kernel, *second = 0; partition path: scatter, count). */
KC_API int kc_ctx_set_timing(kc_ctx* ctx, int enabled);
KC_API int kc_ctx_pass_times(kc_ctx* ctx, float* first_ms, float* second_ms);

/* Device / pinned-host memory helpers so a C caller needs no CUDA headers.
 * (Replaces cudaMallocManaged at main.cu:223,235,251,460,532.)            */
KC_API int kc_device_alloc(kc_ctx* ctx, size_t nbytes, void** d_out);
KC_API int kc_device_free(kc_ctx* ctx, void* d_ptr);
KC_API int kc_host_alloc_pinned(kc_ctx* ctx, size_t nbytes, void** h_out);
KC_API int kc_host_free_pinned(kc_ctx* ctx, void* h_ptr);
KC_API int kc_memcpy_h2d(kc_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes);
KC_API int kc_memcpy_d2h(kc_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes);
KC_API int kc_memset_d(kc_ctx* ctx, void* d_dst, int byte, size_t nbytes);

/* ------------------------------------------------------------------ */
/* k selection / k-mer enumeration.                                    */
/* Replaces compile-time K, PERMS_KMERES (kernels.h:11-19), the        */
/* permutation() odometer (utils.h:21-50), permutationsMap             */
/* (main.cu:134-135) and the c_perms upload (main.cu:151-158).         */
/* ------------------------------------------------------------------ */
#define KC_MAX_K 31          /* uint64 codes */
#define KC_MAX_DENSE_K 16    /* uint32 codes; table = 4^k uint32 must fit in HBM */

/* 4^k, or 0 when k is outside 1..KC_MAX_K */
KC_API uint64_t kc_num_kmers(int k);
/* Same contract as utils.h:21 `permutation(alphabet, length, permutations)`:
 * caller supplies |alphabet|^k buffers of >= k+1 bytes; filled in the
 * reference's little-endian odometer order (entry 1 of "ACGT",3 is "CAA").
 * Unlike the reference the strings ARE NUL-terminated.                     */
KC_API int kc_permutation(const char* alphabet, int k, char** permutations);
/* LE index of a k-mer string (KC_ERR_INVALID if a byte is not in ACGT)     */
KC_API int kc_kmer_index(const char* kmer, int k, uint64_t* idx_out);
/* inverse; `out` needs k+1 bytes                                          */
KC_API int kc_kmer_string(uint64_t idx, int k, char* out);

/* ------------------------------------------------------------------ */
/* Sequence load.  Replaces importSeqs (main.cu:474-545, mode 0) and   */
/* importSeqsNoNL (main.cu:401-473, mode 1) and their globals          */
/* ids / seqs / indexes_aux / data / numberOfSequenses / size_all_seqs */
/* (main.cu:34-35,65-70).  Differences, all deliberate: */
/* terminal offset always emitted; 64-bit offsets; max_seqs<=0 means   */
/* unlimited (reference: MAX_SEQS 100, main.cu:30).                    */
/* ------------------------------------------------------------------ */
#define KC_IMPORT_BLANKLINE 0   /* importSeqs: records end at a blank / CR line */
#define KC_IMPORT_NONL 1        /* importSeqsNoNL: ... or at the next '>' line   */

KC_API int kc_import_seqs(const char* path, int mode, long max_seqs, kc_seqset** out);
/* same parser over an in-memory FASTA image */
KC_API int kc_import_seqs_mem(const char* fasta, size_t nbytes, int mode, long max_seqs,
                              kc_seqset** out);
/* same result as kc_import_seqs_mem with max_seqs <= 0, parsed by `nthreads` host threads
 * (<= 0: one per core, at least 4 MiB of text each).  kc_import_seqs[_mem] use it by
 * themselves for inputs of 32 MiB and more ("next" row f2: ingest at speed).           */
KC_API int kc_import_seqs_mem_threads(const char* fasta, size_t nbytes, int mode, int nthreads,
                                      kc_seqset** out);
/* Device-side form (f2): d_raw = the FASTA file image in DEVICE memory, parsed by three kernels (a
 * seven-state byte transducer whose per-tile maps are composed by a scan); the set's device copies
 * (kc_seqset_to_device) are in place on return, its host image is fetched on the first
 * kc_seqset_data().  h_raw = the same bytes on the host, used only to cut the id strings (NULL: the
 * ids stay empty).  Same result as kc_import_seqs_mem with max_seqs <= 0.                         */
KC_API int kc_import_seqs_device(kc_ctx* ctx, const char* d_raw, const char* h_raw, uint64_t nbytes,
                                 int mode, kc_seqset** out);
/* file form: map the file, copy it to the device as it is, parse it there */
KC_API int kc_import_seqs_gpu(kc_ctx* ctx, const char* path, int mode, kc_seqset** out);
KC_API void kc_seqset_free(kc_seqset* s);
KC_API uint32_t kc_seqset_num_seqs(const kc_seqset* s);
KC_API uint32_t kc_seqset_num_ids(const kc_seqset* s);
/* total bytes of `data` = sum(L_i + 1)  (reference size_all_seqs, main.cu:530) */
KC_API uint64_t kc_seqset_nbytes(const kc_seqset* s);
/* host image of the reference's `data`: sequences back to back, each followed
 * by one '\0' (main.cu:537-543)                                             */
KC_API const char* kc_seqset_data(const kc_seqset* s);
/* num_seqs+1 start offsets into data, last = nbytes (reference `indexes`)     */
KC_API const int64_t* kc_seqset_offsets(const kc_seqset* s);
/* header line i (with its leading '>'), as stored in the reference's `ids`    */
KC_API const char* kc_seqset_id(const kc_seqset* s, uint32_t i);
/* copy data + offsets to the ctx's GPU (idempotent); pointers stay owned by s */
KC_API int kc_seqset_to_device(kc_ctx* ctx, kc_seqset* s, const char** d_data,
                               const int64_t** d_offsets);

/* ------------------------------------------------------------------ */
/* Counting — the hot path.                                            */
/* ------------------------------------------------------------------ */

/* Reference-shaped per-sequence table.  Replaces the kernel
 *   sumKmereCoincidencesGlobalMemory(char* data,int* indices,unsigned num_seqs,int* sum)
 * (kernels.h:113-144, launched main.cu:290): same `data` layout (each sequence
 * followed by one separator byte), same kmer-major result
 *   d_sums[entry + num_seqs * kmer]   (kernels.h:142),
 * but any 1 <= k <= KC_MAX_DENSE_K and 64-bit offsets.  d_sums is overwritten
 * (the reference pre-zeroes it on the host, main.cu:240-242).               */
KC_API int kc_count_per_seq(kc_ctx* ctx, const char* d_data, const int64_t* d_offsets,
                            uint32_t num_seqs, int k, int32_t* d_sums);
KC_API int kc_count_per_seq_async(kc_ctx* ctx, const char* d_data, const int64_t* d_offsets,
                                  uint32_t num_seqs, int k, int32_t* d_sums, void* stream);

/* Aggregate dense table over a byte stream: d_table[idx] = number of valid
 * windows with LE index idx among windows starting in [0, nbytes-k].  Any byte
 * outside ACGT (separator '\0', '\n', 'N', lower case...) resets the window,
 * so separators need no offsets.  Equals the row sums of kc_count_per_seq.
 * d_table (uint32[4^k]) is overwritten.  Counters wrap modulo 2^32.         */
KC_API int kc_count_dense(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k,
                          uint32_t* d_table);
KC_API int kc_count_dense_async(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k,
                                uint32_t* d_table, void* stream);
/* Shard form used by the multi-GPU path: ADDS (no zeroing) the windows that
 * START in [win_begin, win_end) of a buffer whose bytes [0, nbytes) are
 * readable (so the caller passes its shard plus a (k-1)-byte halo).         */
#define KC_DENSE_AUTO 0      /* engine picks by k and size                    */
#define KC_DENSE_DIRECT 1    /* smem-privatised (small k) / global-atomic bins */
#define KC_DENSE_PARTITION 2 /* two-pass radix partition + smem sub-tables     */
/* KC_DENSE_AUTO picks by measurement on B200: k <= 7 shared-memory bins, k = 8
 * KC_DENSE_SMEM16C, k = 12 KC_DENSE_PARTITION_WIDE2 from 2^26 windows, k = 9..11 KC_DENSE_PARTITION from 2^26
 * windows, global REDs otherwise.  The explicit values below exist so that the choices can be compared:        */
#define KC_DENSE_SMEM16C 3   /* k = 8: non-returning shared adds + per-CTA checksum + repair   */
#define KC_DENSE_PARTITION_DEFER 4 /* partition path, full-bin records retried before REDs     */
#define KC_DENSE_PARTITION_PAIR 5  /* k = 12: pass 2 counts two 13-mers + one 12-mer per record
                                      (3 shared increments instead of 5) and folds at the flush */
#define KC_DENSE_PARTITION_WIDE 7  /* k = 12: SEVEN windows per record (the key bits are not stored, so 18
                                      bases still fit a 32-bit slab record): 29 % fewer records through
                                      pass 1; pass 2 counts two 14-mers (4-bit fields) + one 12-mer   */
#define KC_DENSE_PARTITION_TRIO 6  /* k = 12: one 14-mer (8-bit fields) + one 13-mer per record:
                                      2 shared increments instead of 5                           */
#define KC_DENSE_PARTITION_DEFER_PAIR 8  /* k = 12: the scatter of 4 with the count of 5 */
#define KC_DENSE_PARTITION_DEFER_TRIO 9  /* k = 12: the scatter of 4 with the count of 6 */
#define KC_DENSE_PARTITION_WIDE2 10  /* k = 12: seven windows per record, second-generation scatter: super-steps of
                                        7 x 512 bytes = 16 records per lane, bin generation = write cursor, flush by
                                        the lane that fills the bin; the KC_DENSE_AUTO choice       */
KC_API int kc_count_dense_range_async(kc_ctx* ctx, const char* d_data, uint64_t nbytes,
                                      uint64_t win_begin, uint64_t win_end, int k,
                                      uint32_t* d_table, int algo, void* stream);
/* End to end from HOST memory: chunked H2D (pinned staging, double buffered)
 * overlapped with counting, table copied back to h_table.                   */
KC_API int kc_count_dense_host(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k,
                               uint32_t* h_table);

/* Sparse counting for k <= 31 (uint64 LE codes).  Result: distinct k-mers
 * sorted by code with their uint32 counts.                                  */
#define KC_SPARSE_HASH 0   /* open-addressing hash table (CAS key, RED count) */
#define KC_SPARSE_SORT 1   /* radix sort of codes + run-length reduce          */
#define KC_SPARSE_RADIX 2  /* MSD radix partition, leaves sorted in shared memory (k >= 11);
                              falls back to KC_SPARSE_HASH when skewed data overflows a region, and
                              takes that path directly for inputs below 4 M windows            */
#define KC_SPARSE_AUTO 3   /* the engine picks: KC_SPARSE_RADIX wherever it exists (k >= 11), chosen by
                              measurement on B200, else KC_SPARSE_HASH          */
/* OR-ed into `algo`: leave the distinct (code,count) pairs in table order instead of
 * sorting them — for callers that re-bucket and merge anyway (the multi-GPU path).   */
#define KC_SPARSE_UNSORTED 0x100
/* OR-ed into KC_SPARSE_RADIX: return KC_ERR_TABLE_FULL instead of recounting with the hash path */
#define KC_SPARSE_NO_FALLBACK 0x200
KC_API int kc_count_sparse(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, int algo,
                           uint64_t capacity_hint, kc_sparse** out);
/* The stages of KC_SPARSE_RADIX, exposed for the multi-GPU path ("each GPU owns a disjoint key
 * range"): every rank scatters its reads into level-1 partitions (the top 10 bits of the code),
 * the ranks exchange the slab blocks of each other's partition range with one all-to-all (the
 * slabs are partition-major, so a range is one contiguous block), and every rank counts the
 * partitions it owns.  Concatenating the ranks' results in rank order gives the sorted whole. */
typedef struct kc_radix_plan {
    int32_t k;
    uint32_t world;           /* ranks the partitions are split over; must divide `partitions` */
    uint32_t partitions;      /* level-1 partitions (1024)                                      */
    uint32_t parts_per_rank;  /* partitions / world                                             */
    uint32_t grid;            /* pass-1 CTAs = regions per partition and rank                   */
    uint32_t rec_bytes;       /* size of a level-1 record: 4 (k <= 21) or 8                     */
    uint32_t shape;           /* 0 = shipped 1024 x 1024, 1 = 16 x 16 (KC_SPARSE_RADIX_SHAPE=small, tests) */
    uint32_t round_bits;      /* the count runs in 2^round_bits ROUNDS: round r holds the windows whose top round_bits
                                 code bits are r (chosen by kc_sparse_radix_plan so that a leaf fits shared memory and
                                 the slabs fit the device; 0 = one round)                        */
    uint64_t max_windows;     /* the plan is good for inputs of up to this many windows per rank */
    uint64_t region_records;  /* capacity of one (partition, CTA) region                         */
    uint64_t slab_bytes;      /* partitions * grid * region_records * rec_bytes                  */
    uint64_t counts_bytes;    /* partitions * grid * 4                                           */
} kc_radix_plan;
/* Give back what the ctx keeps between calls: its scratch areas (partition slabs, staging buffers) and the idle blocks
 * of the device's memory pool.  They grow again on demand; results (kc_sparse) and seqsets are not touched.          */
KC_API void kc_ctx_release_memory(kc_ctx* ctx);
/* Device memory the caller's own allocator has cached and will reuse for the buffers it passes in (a torch caller:
 * memory_reserved - memory_allocated): kc_sparse_radix_plan counts it as available when it sizes the rounds. */
KC_API void kc_ctx_set_reusable_bytes(kc_ctx* ctx, uint64_t nbytes);
/* same plan on every rank: pass the LARGEST per-rank window count */
KC_API int kc_sparse_radix_plan(kc_ctx* ctx, uint64_t max_windows_per_rank, int k, uint32_t world,
                                kc_radix_plan* plan);
/* the same with at least 2^min_round_bits rounds: what a caller asks for after KC_ERR_TABLE_FULL from a count whose
 * leaves turned out denser than planned (kc_count_sparse and the sharded host logic retry twice this way)           */
KC_API int kc_sparse_radix_plan_rounds(kc_ctx* ctx, uint64_t max_windows_per_rank, int k, uint32_t world,
                                       uint32_t min_round_bits, kc_radix_plan* plan);
/* d_slabs: plan->slab_bytes, d_counts: plan->counts_bytes, both [partition][cta]...; synchronous;
 * KC_ERR_TABLE_FULL when a region overflowed (skewed input)                                  */
KC_API int kc_sparse_radix_scatter(kc_ctx* ctx, const char* d_data, uint64_t nbytes,
                                   const kc_radix_plan* plan, void* d_slabs, uint32_t* d_counts);
/* counts `nparts` partitions starting at global partition `part_first`; d_slabs / d_counts hold
 * `nsrc` blocks of [nparts][grid][region_records] / [nparts][grid], one per source rank (what an
 * equal-split all-to-all of the scatter outputs delivers; nsrc = 1, nparts = partitions on one GPU) */
KC_API int kc_sparse_radix_count(kc_ctx* ctx, const kc_radix_plan* plan, const void* d_slabs,
                                 const uint32_t* d_counts, uint32_t nsrc, uint32_t part_first,
                                 uint32_t nparts, kc_sparse** out);
/* The same two stages for ONE round of a plan with round_bits > 0 (run the rounds 0 .. 2^round_bits - 1
 * one after the other, each with its own all-to-all; the slab buffers can be reused).  Rank r's results of
 * consecutive rounds are ascending code ranges: kc_sparse_concat joins them (the pieces stay valid).    */
KC_API int kc_sparse_radix_scatter_round(kc_ctx* ctx, const char* d_data, uint64_t nbytes,
                                         const kc_radix_plan* plan, uint32_t round, void* d_slabs,
                                         uint32_t* d_counts);
KC_API int kc_sparse_radix_count_round(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round,
                                       const void* d_slabs, const uint32_t* d_counts, uint32_t nsrc,
                                       uint32_t part_first, uint32_t nparts, kc_sparse** out);
/* The count stage of round `round` APPENDED to one result: call it for the rounds in ascending order with *acc = NULL
 * before the first; after every call *acc holds all k-mers counted so far (ascending), after the last one it is the
 * rank's whole result.  The arrays are sized from the first round (count x rounds + 12.5 %) and grow when a later
 * round does not fit, so nothing is concatenated and no second copy of the result exists (kc_sparse_radix_count_round
 * + kc_sparse_concat: pieces and whole side by side).  On an error *acc stays valid (without the failed round);
 * free it with kc_sparse_free. */
KC_API int kc_sparse_radix_count_round_append(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round,
                                              const void* d_slabs, const uint32_t* d_counts, uint32_t nsrc,
                                              uint32_t part_first, uint32_t nparts, kc_sparse** acc);
KC_API int kc_sparse_concat(kc_ctx* ctx, kc_sparse* const* parts, uint32_t nparts, kc_sparse** out);
KC_API void kc_sparse_free(kc_sparse* s);
KC_API uint64_t kc_sparse_size(const kc_sparse* s);
KC_API const uint64_t* kc_sparse_d_keys(const kc_sparse* s);    /* device, sorted */
KC_API const uint32_t* kc_sparse_d_counts(const kc_sparse* s);  /* device */
KC_API int kc_sparse_copy_to_host(kc_ctx* ctx, const kc_sparse* s, uint64_t* h_keys,
                                  uint32_t* h_counts);
/* Building blocks of the hash-sharded multi-GPU path (owner = mix64(code) % G):
 * split a (keys,counts) list into G owner buckets (stable within a bucket),   */
KC_API int kc_sparse_bucket_by_owner(kc_ctx* ctx, const uint64_t* d_keys,
                                     const uint32_t* d_counts, uint64_t n, uint32_t num_owners,
                                     uint64_t* d_keys_out, uint32_t* d_counts_out,
                                     uint64_t* h_bucket_sizes /* [num_owners] */);
/* merge an unsorted (keys,counts) list with duplicates into a kc_sparse       */
KC_API int kc_sparse_merge(kc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_counts,
                           uint64_t n, kc_sparse** out);
/* owner hash, exported so host-side sharding logic and tests agree with the GPU */
KC_API uint64_t kc_mix64(uint64_t code);

/* Full-scale self-checks of a count (csrc/check.cu).  The reference verifies itself by running its CPU
 * path beside its GPU path on the same input (main.cu:169,172) and diffing the outputs by hand; at the
 * sizes of BASELINE configs 3-5 the equivalent is a multiset fingerprint with h = kc_mix64:
 *   kc_window_fingerprint: F = sum over the VALID windows of the input of h(code), and their number,
 *                          from one streaming scan (additive over shards and ranks);
 *   kc_sparse_fingerprint / kc_dense_fingerprint: F = sum over the k-mers of count * h(code), and the
 *                          sum of the counts, from a result.
 * A correct count has equal F (mod 2^64) and equal totals.  Results are written to HOST pointers. */
KC_API int kc_window_fingerprint(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k,
                                 uint64_t* h_fp, uint64_t* h_windows);
/* *h_descents (may be NULL): positions where the keys are not strictly ascending — 0 for a valid result */
KC_API int kc_sparse_fingerprint(kc_ctx* ctx, const kc_sparse* s, uint64_t* h_fp, uint64_t* h_total,
                                 uint64_t* h_descents);
KC_API int kc_dense_fingerprint(kc_ctx* ctx, const uint32_t* d_table, int k, uint64_t* h_fp,
                                uint64_t* h_total);

/* ------------------------------------------------------------------ */
/* "Next" row f4: the 2-bit packed sequence store sketched in the      */
/* reference's comments (main.cu:78-86, utils.h:65-92: "AACG ->        */
/* 00000110", four bases per byte, the first base in the two most      */
/* significant bits, A=00 C=01 G=10 T=11), plus a validity bitmap so   */
/* that N / separators / lower case keep resetting the window:         */
/* bit i%32 of word i/32 (LSB first) is set iff byte i was not an      */
/* upper-case ACGT; such bases pack as 00.  0.375 B/base at rest.      */
/* ------------------------------------------------------------------ */
KC_API uint64_t kc_packed_bytes(uint64_t nbases);   /* (n+3)/4        */
KC_API uint64_t kc_badmask_bytes(uint64_t nbases);  /* (n+31)/32 * 4  */
/* d_data 16-byte aligned (any cudaMalloc'ed buffer), d_packed 4-byte aligned */
KC_API int kc_pack_2bit(kc_ctx* ctx, const char* d_data, uint64_t nbytes, void* d_packed,
                        uint32_t* d_badmask, void* stream);
/* inverse; invalid bases come back as 'N' (their identity is not stored)     */
KC_API int kc_unpack_2bit(kc_ctx* ctx, const void* d_packed, const uint32_t* d_badmask,
                          uint64_t nbases, char* d_data_out, void* stream);
/* == kc_count_dense of the bytes the store was packed from (synchronous; d_table overwritten):
 * 2^30-base chunks are unpacked into an ASCII scratch and counted by the ordinary dense path */
KC_API int kc_count_dense_packed(kc_ctx* ctx, const void* d_packed, const uint32_t* d_badmask,
                                 uint64_t nbases, int k, uint32_t* d_table);
/* The same conversion on the HOST cores (format conversion only; counting has no CPU path): h_data ->
 * h_packed[(n+3)/4], h_badmask[(n+31)/32]; bit-identical to kc_pack_2bit's output.  nthreads 0 = one per
 * core this process may run on (at most 64).  AVX-512BW or AVX2 body when the host has it (kc_host_pack_simd()). */
KC_API int kc_pack_2bit_host(const char* h_data, uint64_t nbytes, void* h_packed, uint32_t* h_badmask,
                             int nthreads);
KC_API int kc_host_pack_threads(int nthreads);   /* packer threads a call with this `nthreads` argument uses
                                                    (0: cores the process may run on - 1, cgroup quota, at most 64) */
KC_API int kc_host_pack_simd(void);   /* body kc_pack_2bit_host uses here: 0 scalar, 1 AVX2, 2 AVX-512BW */
/* diagnostic / test aid: the same with a chosen loop body (0 = best available, 1 = scalar, 2 = AVX2 if present) */
KC_API int kc_pack_2bit_host_body(const char* h_data, uint64_t nbytes, void* h_packed,
                                  uint32_t* h_badmask, int nthreads, int body);
/* kc_count_dense_host (main.cu:287-299: the reference's count step starts from host/managed memory)
 * for hosts with cores to spare: packer threads turn the ASCII into the store's layout slot by slot
 * (pinned ring), the slots cross PCIe at 0.25 bytes per base + the 4 KiB bitmap blocks that hold an invalid
 * byte (at most 0.375 in all) instead of 1, the GPU rebuilds bitmap and bytes at HBM speed and counts behind
 * the copies.  h_data need not be pinned (the packer threads read it; only the library's own ring is).
 * Same result as kc_count_dense_host / kc_count_dense.  nthreads =
 * packer threads (0 = cores available to the process - 1, at most 64).                               */
KC_API int kc_count_dense_host_packed(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k,
                                      uint32_t* h_table, int nthreads);
/* the same, counted into the caller's DEVICE table (overwritten; synchronous) — for callers that go on
 * with the table on the GPU: the multi-GPU reduce of a rank's shard, the distance step                */
KC_API int kc_count_dense_host_packed_dev(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k,
                                          uint32_t* d_table, int nthreads);

/* ------------------------------------------------------------------ */
/* Count table dump.  Byte-identical to the (commented-out) dump at    */
/* main.cu:301-309: "Sums:\n", per k-mer row "%d: " then "%d,\t" per   */
/* sequence then "\n", and a final "\n".  h_sums is a HOST table       */
/* [4^k][num_seqs].  path == NULL writes to stdout.                    */
/* ------------------------------------------------------------------ */
KC_API int kc_dump_counts(const char* path, const int32_t* h_sums, int k, uint32_t num_seqs);

/* ------------------------------------------------------------------ */
/* "Next" row f1: k-mer distance step.  Replaces the num_seqs          */
/* synchronous launches of minKmeres2 (kernels.h:85-109, main.cu:327-  */
/* 335) with one fused all-pairs kernel.  d_dist is the reference's    */
/* packed strict upper triangle (kernels.h:46-48), float32,            */
/* n(n-1)/2 entries: d = 1 - sum_kmer min(c_i,c_j)/(min(L_i,L_j)-k+1). */
/* ------------------------------------------------------------------ */
KC_API int kc_kmer_distance(kc_ctx* ctx, const int32_t* d_sums, const int64_t* d_offsets,
                            uint32_t num_seqs, int k, float* d_dist);
KC_API int64_t kc_triangular_index(int64_t i, int64_t j, int64_t n); /* kernels.h:46-48 */
/* one "%f\n" per pair (main.cu:355-358); path == NULL writes to stdout */
KC_API int kc_dump_distances(const char* path, const float* h_dist, uint64_t n_pairs);

/* ------------------------------------------------------------------ */
/* Deterministic synthetic inputs, generated directly in  */
/* HBM so 3.1-30 Gbp configs never cross PCIe.  The oracle has the     */
/* same generators on the CPU (oracle/kmer_oracle.c).                  */
/* ------------------------------------------------------------------ */
/* base(i) = "ACGT"[splitmix64(seed + pos0 + i) >> 62], i in [0,n)            */
KC_API int kc_gen_bases(kc_ctx* ctx, uint64_t seed, uint64_t pos0, uint64_t n, char* d_out,
                        void* stream);
/* config-3 style genome slice [pos0,pos0+n) of a total_len sequence: random
 * bases with `long_runs` long N runs and `short_runs` short ones overlaid    */
KC_API int kc_gen_genome(kc_ctx* ctx, uint64_t seed, uint64_t total_len, uint32_t long_runs,
                         uint32_t short_runs, int k, uint64_t pos0, uint64_t n, char* d_out,
                         void* stream);
/* config-4/5 style reads [read0, read0+nreads): read_len bases sampled from a
 * genome_len random genome with 1/err_den substitutions, each followed by '\n' */
KC_API int kc_gen_reads(kc_ctx* ctx, uint64_t seed, uint64_t genome_len, uint32_t read_len,
                        uint32_t err_den, uint64_t read0, uint64_t nreads, char* d_out,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KMER_B200_H */

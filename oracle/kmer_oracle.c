/*
 * kmer_oracle.c — CPU restatement of the k-mer counting path of
 * axlwild/dna-kmeres-parallel.   *** TEST INFRASTRUCTURE ONLY ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline.  The product (libkmerb200.so) never links,
 * loads or calls it; there is no CPU fallback in the product.
 *
 * Parity pin: this restatement is checked (tests/test_oracle.py) against
 *   (a) the reference's OWN permutation() + permutationsCountAll() +
 *       importSeqs() + sequentialKmerCount2(), compiled unmodified from
 *       /root/reference by oracle/build_ref.sh into oracle/_ref/ (k = 3..6), and
 *   (b) the golden vectors of SURVEY.md §8c / tests/golden/ (JSON fixtures) that were
 *       generated from (a) by tests/golden/make_golden.py.
 * For k > 6 (the reference cannot be compiled: kernels.h:21 overflows constant
 * memory) the same code path is used with a larger k; nothing else changes.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OR_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* alphabet: "ACGT" in this order (main.cu:122) -> digit 0..3; every   */
/* other byte (N, lower case, '\r', '\0', '|') is invalid: the         */
/* reference's map lookup finds no entry for a window holding one      */
/* (main.cu:643-644) and its GPU compare matches nothing               */
/* (kernels.h:136-140).                                                */
/* ------------------------------------------------------------------ */
static inline int or_code(unsigned char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}

OR_API uint64_t or_num_kmers(int k) { return (k < 1 || k > 31) ? 0 : (1ull << (2 * k)); }

/* utils.h:21-50 — odometer whose digit 0 (string position 0) increments first,
 * i.e. entry i spells the base-|alphabet| digits of i, least significant first. */
OR_API void or_permutation(const char* alphabet, int k, char** out) {
    size_t na = strlen(alphabet);
    uint64_t total = 1;
    for (int i = 0; i < k; i++) total *= na;
    for (uint64_t i = 0; i < total; i++) {
        uint64_t v = i;
        for (int p = 0; p < k; p++) {
            out[i][p] = alphabet[v % na];
            v /= na;
        }
        out[i][k] = '\0';
    }
}

/* main.cu:134-135: permutationsMap[perms[i]] = i+1  =>  0-based LE index */
OR_API int or_kmer_index(const char* s, int k, uint64_t* idx) {
    uint64_t v = 0;
    for (int p = 0; p < k; p++) {
        int c = or_code((unsigned char)s[p]);
        if (c < 0) return -1;
        v |= (uint64_t)c << (2 * p);
    }
    *idx = v;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Counting.                                                           */
/* ------------------------------------------------------------------ */

/* Naive form, as close to main.cu:636-646 as C allows: for every window
 * i in [0, L-k] take the k bytes, look the string up, bucket 0 = "not a
 * key" (main.cu:643-644), bucket idx+1 otherwise (main.cu:134-135).
 * `seq` is WITHOUT the trailing '|' the reference appends (main.cu:505):
 * its loop bound `sequence_len - k` with the '|' included is L-k+1 windows.
 * counts has 4^k + 1 entries.  Used to cross-check the rolling form.     */
OR_API void or_count_all_naive(const char* seq, uint64_t len, int k, int64_t* counts) {
    uint64_t nk = or_num_kmers(k);
    for (uint64_t i = 0; i <= nk; i++) counts[i] = 0;
    if (len < (uint64_t)k) return;
    for (uint64_t i = 0; i + k <= len; i++) {
        uint64_t idx;
        if (or_kmer_index(seq + i, k, &idx) == 0)
            counts[idx + 1]++;
        else
            counts[0]++;
    }
}

/* Rolling form of the same thing over an arbitrary byte stream: windows
 * starting in [win_begin, win_end) of data[0..n).  A window is valid iff
 * its k bytes are all ACGT, so separators ('\0' of main.cu:537-543, '\n')
 * split sequences without offsets (no window spans two sequences,
 * main.cu:600-603 loops per sequence).  table[idx] += 1 per valid window
 * (GPU row = idx, kernels.h:117,142); *invalid += 1 per invalid one
 * (CPU bucket 0).  Counters are uint32 and wrap (documented engine
 * behaviour; the reference's int would be UB).                         */
OR_API void or_count_dense_range(const unsigned char* data, uint64_t n, uint64_t win_begin,
                                 uint64_t win_end, int k, uint32_t* table, uint64_t* invalid) {
    if (n < (uint64_t)k) return;
    uint64_t last = n - k + 1; /* number of windows in the stream */
    if (win_end > last) win_end = last;
    if (win_begin >= win_end) return;
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    const int top = 2 * (k - 1);
    uint64_t code = 0;
    int run = 0; /* valid bases seen in a row, saturating at k */
    uint64_t inv = 0;
    /* prime with the k-1 bytes before the first window's last byte */
    for (uint64_t q = win_begin; q < win_end + k - 1; q++) {
        int c = or_code(data[q]);
        if (c < 0) {
            run = 0;
            code = 0;
        } else {
            code = ((code >> 2) | ((uint64_t)c << top)) & mask;
            if (run < k) run++;
        }
        if (q + 1 >= win_begin + k) { /* window starting at q-k+1 is complete */
            if (run >= k)
                table[code]++;
            else
                inv++;
        }
    }
    if (invalid) *invalid += inv;
}

OR_API void or_count_dense(const unsigned char* data, uint64_t n, int k, uint32_t* table,
                           uint64_t* invalid) {
    uint64_t nk = or_num_kmers(k);
    memset(table, 0, nk * sizeof(uint32_t));
    if (invalid) *invalid = 0;
    or_count_dense_range(data, n, 0, n, k, table, invalid);
}

/* All host cores: chunk + (k-1) halo, per-thread tables merged. */
typedef struct {
    const unsigned char* data;
    uint64_t n, b, e;
    int k;
    uint32_t* table;
    uint64_t invalid;
} or_mt_job;

static void* or_mt_worker(void* p) {
    or_mt_job* j = (or_mt_job*)p;
    or_count_dense_range(j->data, j->n, j->b, j->e, j->k, j->table, &j->invalid);
    return NULL;
}

OR_API int or_count_dense_mt(const unsigned char* data, uint64_t n, int k, uint32_t* table,
                             uint64_t* invalid, int nthreads) {
    uint64_t nk = or_num_kmers(k);
    memset(table, 0, nk * sizeof(uint32_t));
    if (invalid) *invalid = 0;
    if (nthreads < 1) nthreads = 1;
    if (n < (uint64_t)k) return 0;
    uint64_t nwin = n - k + 1;
    or_mt_job* jobs = (or_mt_job*)calloc(nthreads, sizeof(or_mt_job));
    pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
    if (!jobs || !th) return -1;
    for (int t = 0; t < nthreads; t++) {
        jobs[t].data = data;
        jobs[t].n = n;
        jobs[t].k = k;
        jobs[t].b = nwin * (uint64_t)t / nthreads;
        jobs[t].e = nwin * (uint64_t)(t + 1) / nthreads;
        jobs[t].table = (t == 0) ? table : (uint32_t*)calloc(nk, sizeof(uint32_t));
        if (!jobs[t].table) return -1;
        pthread_create(&th[t], NULL, or_mt_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (invalid) *invalid += jobs[t].invalid;
        if (t > 0) {
            for (uint64_t i = 0; i < nk; i++) table[i] += jobs[t].table[i];
            free(jobs[t].table);
        }
    }
    free(jobs);
    free(th);
    return 0;
}

/* Reference-shaped per-sequence table.  Input layout = main.cu:537-543
 * (`data`: each sequence followed by one separator byte) + offsets with
 * num_seqs+1 entries (main.cu:519-523); L_e = off[e+1]-off[e]-1
 * (kernels.h:124 counts the separator).  Output kmer-major
 * sums[e + num_seqs*idx] (kernels.h:142); invalid[e] = CPU bucket 0.    */
OR_API void or_count_per_seq(const unsigned char* data, const int64_t* offsets, uint32_t num_seqs,
                             int k, int32_t* sums, uint64_t* invalid) {
    uint64_t nk = or_num_kmers(k);
    memset(sums, 0, nk * (uint64_t)num_seqs * sizeof(int32_t));
    uint32_t* tmp = (uint32_t*)malloc(nk * sizeof(uint32_t));
    for (uint32_t e = 0; e < num_seqs; e++) {
        int64_t L = offsets[e + 1] - offsets[e] - 1;
        uint64_t inv = 0;
        memset(tmp, 0, nk * sizeof(uint32_t));
        if (L >= k) or_count_dense_range(data + offsets[e], (uint64_t)L, 0, (uint64_t)L, k, tmp, &inv);
        for (uint64_t i = 0; i < nk; i++) sums[e + (uint64_t)num_seqs * i] = (int32_t)tmp[i];
        if (invalid) invalid[e] = inv;
    }
    free(tmp);
}

/* Sparse form for k <= 31: all valid window codes, sorted, run-length
 * reduced.  Returns the number of distinct k-mers; *keys / *counts are
 * malloc'd (free with or_free).                                          */
static int or_cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

static void or_radix_sort_u64(uint64_t* a, uint64_t n, int bits) {
    if (n < 4096) {
        qsort(a, n, sizeof(uint64_t), or_cmp_u64);
        return;
    }
    uint64_t* b = (uint64_t*)malloc(n * sizeof(uint64_t));
    uint64_t* src = a;
    uint64_t* dst = b;
    for (int shift = 0; shift < bits; shift += 11) {
        uint64_t cnt[2049];
        memset(cnt, 0, sizeof(cnt));
        for (uint64_t i = 0; i < n; i++) cnt[((src[i] >> shift) & 2047) + 1]++;
        for (int i = 0; i < 2048; i++) cnt[i + 1] += cnt[i];
        for (uint64_t i = 0; i < n; i++) dst[cnt[(src[i] >> shift) & 2047]++] = src[i];
        uint64_t* t = src;
        src = dst;
        dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(b);
}

OR_API int64_t or_count_sparse(const unsigned char* data, uint64_t n, int k, uint64_t** keys,
                               uint32_t** counts, uint64_t* invalid) {
    *keys = NULL;
    *counts = NULL;
    if (invalid) *invalid = 0;
    if (k < 1 || k > 31 || n < (uint64_t)k) return 0;
    uint64_t nwin = n - k + 1;
    uint64_t* codes = (uint64_t*)malloc(nwin * sizeof(uint64_t));
    if (!codes) return -1;
    const uint64_t mask = (1ull << (2 * k)) - 1;
    const int top = 2 * (k - 1);
    uint64_t code = 0, m = 0, inv = 0;
    int run = 0;
    for (uint64_t q = 0; q < n; q++) {
        int c = or_code(data[q]);
        if (c < 0) {
            run = 0;
            code = 0;
        } else {
            code = ((code >> 2) | ((uint64_t)c << top)) & mask;
            if (run < k) run++;
        }
        if (q + 1 >= (uint64_t)k) {
            if (run >= k)
                codes[m++] = code;
            else
                inv++;
        }
    }
    if (invalid) *invalid = inv;
    or_radix_sort_u64(codes, m, 2 * k);
    uint64_t d = 0;
    for (uint64_t i = 0; i < m; i++)
        if (i == 0 || codes[i] != codes[i - 1]) d++;
    uint64_t* K = (uint64_t*)malloc((d ? d : 1) * sizeof(uint64_t));
    uint32_t* C = (uint32_t*)malloc((d ? d : 1) * sizeof(uint32_t));
    uint64_t w = 0;
    for (uint64_t i = 0; i < m; i++) {
        if (i == 0 || codes[i] != codes[i - 1]) {
            K[w] = codes[i];
            C[w] = 1;
            w++;
        } else {
            C[w - 1]++;
        }
    }
    free(codes);
    *keys = K;
    *counts = C;
    return (int64_t)d;
}

OR_API void or_free(void* p) { free(p); }

/* ------------------------------------------------------------------ */
/* Distance step ("next" row f1): sequentialKmerCount2 main.cu:587-621. */
/* sums is the kmer-major table above; L_e from offsets.  Sum of mins   */
/* in integer (the CPU uses long, main.cu:593,607-613), then            */
/* 1 - (float)sum / (minLength - k + 1)  (main.cu:614), stored at the   */
/* packed index of main.cu:671-673 with 1-based i and gap j-i.          */
/* ------------------------------------------------------------------ */
OR_API int64_t or_triangular_index(int64_t i, int64_t j, int64_t n) {
    return (n * (i - 1) - (((i - 2) * (i - 1)) / 2)) + (j - i);
}

OR_API void or_distance(const int32_t* sums, const int64_t* offsets, uint32_t num_seqs, int k,
                        float* dist) {
    uint64_t nk = or_num_kmers(k);
    long n = num_seqs;
    for (long i = 0; i < n - 1; i++) {
        for (long j = i + 1; j < n; j++) {
            long Li = offsets[i + 1] - offsets[i] - 1, Lj = offsets[j + 1] - offsets[j] - 1;
            long minLength = Li < Lj ? Li : Lj;
            long sum = 0;
            for (uint64_t p = 0; p < nk; p++) {
                long a = sums[i + (uint64_t)n * p], b = sums[j + (uint64_t)n * p];
                sum += a < b ? a : b;
            }
            float distance = 1 - (float)sum / (minLength - k + 1);
            dist[or_triangular_index(i + 1, j - i, n)] = distance;
        }
    }
}

/* ------------------------------------------------------------------ */
/* FASTA loader restatement: importSeqs main.cu:474-545 (mode 0) and    */
/* importSeqsNoNL main.cu:401-473 (mode 1), over an in-memory image.    */
/* Mirrors the engine's documented deviations (terminal offset always   */
/* emitted, 64-bit offsets, max_seqs <= 0 unlimited) so it can be       */
/* compared 1:1 with kc_import_seqs; the reference-exact quirks are     */
/* checked against oracle/_ref in tests/test_loader.py.                 */
/* Output: data (sequence bytes each followed by '\0'), offsets[n+1],   */
/* ids concatenated with '\n'.  All malloc'd.                           */
/* ------------------------------------------------------------------ */
typedef struct {
    char* p;
    size_t n, cap;
} or_buf;

static void or_buf_add(or_buf* b, const char* s, size_t n) {
    if (b->n + n + 1 > b->cap) {
        b->cap = (b->n + n + 1) * 2 + 64;
        b->p = (char*)realloc(b->p, b->cap);
    }
    memcpy(b->p + b->n, s, n);
    b->n += n;
    b->p[b->n] = 0;
}

/* std::getline semantics: returns 1 and the line (without '\n') while any
 * characters remain; a final line without '\n' is still returned.        */
static int or_getline(const char* f, size_t n, size_t* pos, const char** line, size_t* len) {
    if (*pos >= n) return 0;
    const char* s = f + *pos;
    const char* e = (const char*)memchr(s, '\n', n - *pos);
    if (e) {
        *len = (size_t)(e - s);
        *pos += *len + 1;
    } else {
        *len = n - *pos;
        *pos = n;
    }
    *line = s;
    return 1;
}

OR_API int or_import_seqs_mem(const char* fasta, size_t nbytes, int mode, long max_seqs,
                              char** data_out, uint64_t* data_len, int64_t** offsets_out,
                              uint32_t* num_seqs, char** ids_out, uint32_t* num_ids) {
    or_buf data = {0}, ids = {0}, acc = {0};
    int64_t* offs = NULL;
    size_t noffs = 0, capoffs = 0;
    uint32_t nseq = 0, nid = 0;
    size_t pos = 0, len;
    const char* line;
    int newSeq = 0, stop = 0;
#define PUSH_OFF(v)                                                       \
    do {                                                                  \
        if (noffs == capoffs) {                                           \
            capoffs = capoffs * 2 + 16;                                   \
            offs = (int64_t*)realloc(offs, capoffs * sizeof(int64_t));    \
        }                                                                 \
        offs[noffs++] = (v);                                              \
    } while (0)
#define FINISH_RECORD()                       \
    do {                                      \
        PUSH_OFF((int64_t)data.n);            \
        for (size_t q_ = 0; q_ < acc.n; q_++) /* main.cu:538-541: every '|' -> NUL */ \
            if (acc.p[q_] == '|') acc.p[q_] = 0; \
        or_buf_add(&data, acc.p, acc.n);      \
        or_buf_add(&data, "\0", 1);           \
        nseq++;                               \
        acc.n = 0;                            \
    } while (0)
    or_buf_add(&acc, "", 0);
    or_buf_add(&data, "", 0);
    or_buf_add(&ids, "", 0);
    while (!stop && or_getline(fasta, nbytes, &pos, &line, &len)) {
        if (len == 0) continue;                       /* main.cu:490-492 */
        if (line[0] == '>') {                         /* main.cu:494-499 */
            or_buf_add(&ids, line, len);
            or_buf_add(&ids, "\n", 1);
            nid++;
            newSeq = 1;
            continue;
        }
        if (!newSeq) continue;                        /* stray line: dropped */
        newSeq = 0;
        acc.n = 0;
        or_buf_add(&acc, line, len);                  /* main.cu:502 */
        int closed = 0;
        while (or_getline(fasta, nbytes, &pos, &line, &len)) {   /* main.cu:503 */
            int gt = (mode == 1 && len > 0 && line[0] == '>');
            if (gt) newSeq = 1;                       /* main.cu:431 */
            if (len == 0 || line[0] == 13 || gt) {    /* main.cu:504 / 432 */
                FINISH_RECORD();
                closed = 1;
                break;
            }
            or_buf_add(&acc, line, len);              /* main.cu:513 */
            if (max_seqs > 0 && (long)nseq >= max_seqs) break;   /* main.cu:514 */
        }
        if (!closed && acc.n > 0) {                   /* main.cu:516-525 */
            FINISH_RECORD();
            if (max_seqs > 0 && (long)nseq >= max_seqs) stop = 1;
        }
    }
    PUSH_OFF((int64_t)data.n); /* terminal offset: always (engine deviation) */
    *data_out = data.p;
    *data_len = data.n;
    *offsets_out = offs;
    *num_seqs = nseq;
    *ids_out = ids.p;
    *num_ids = nid;
    free(acc.p);
    return 0;
#undef PUSH_OFF
#undef FINISH_RECORD
}

/* Count table dump, main.cu:301-309 (commented-out block): text into a
 * malloc'd buffer.                                                       */
OR_API char* or_dump_counts(const int32_t* sums, int k, uint32_t num_seqs) {
    uint64_t nk = or_num_kmers(k);
    or_buf b = {0};
    char tmp[64];
    or_buf_add(&b, "Sums:\n", 6);
    uint64_t idx = 0;
    for (uint64_t j = 0; j < nk; j++) {
        int n = snprintf(tmp, sizeof tmp, "%d: ", (int)j);
        or_buf_add(&b, tmp, n);
        for (uint32_t i = 0; i < num_seqs; i++) {
            n = snprintf(tmp, sizeof tmp, "%d,\t", sums[idx++]);
            or_buf_add(&b, tmp, n);
        }
        or_buf_add(&b, "\n", 1);
    }
    or_buf_add(&b, "\n", 1);
    return b.p;
}

/* ------------------------------------------------------------------ */
/* Deterministic synthetic inputs.  The   */
/* GPU generators in csrc/gen.cu produce identical bytes.               */
/* ------------------------------------------------------------------ */
OR_API uint64_t or_sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* owner hash of the hash-sharded multi-GPU path (same function as kc_mix64) */
OR_API uint64_t or_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33;
    x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}

OR_API void or_gen_bases(uint64_t seed, uint64_t pos0, uint64_t n, char* out) {
    for (uint64_t i = 0; i < n; i++) out[i] = "ACGT"[or_sm64(seed + pos0 + i) >> 62];
}

#define OR_SHORT_TAG 0x5EED000000000000ull

static void or_overlay_run(uint64_t start, uint64_t len, uint64_t total, uint64_t pos0, uint64_t n,
                           char* out) {
    uint64_t end = start + len;
    if (end > total) end = total;
    uint64_t a = start > pos0 ? start : pos0;
    uint64_t b = end < pos0 + n ? end : pos0 + n;
    for (uint64_t p = a; p < b; p++) out[p - pos0] = 'N';
}

OR_API void or_gen_genome(uint64_t seed, uint64_t total_len, uint32_t long_runs,
                          uint32_t short_runs, int k, uint64_t pos0, uint64_t n, char* out) {
    or_gen_bases(seed, pos0, n, out);
    if (long_runs) {
        uint64_t pitch = total_len / long_runs;
        uint64_t jit = pitch / 2;
        if (jit == 0) jit = 1;
        for (uint64_t r = 0; r < long_runs; r++) {
            uint64_t start = r * pitch + or_sm64(seed ^ r) % jit;
            uint64_t len = 1 + or_sm64(seed ^ ~r) % 300000ull;
            or_overlay_run(start, len, total_len, pos0, n, out);
        }
    }
    for (uint64_t j = 0; j < short_runs; j++) {
        uint64_t t = OR_SHORT_TAG + j;
        uint64_t start = or_sm64(seed ^ t) % total_len;
        uint64_t len = 1 + or_sm64(seed ^ ~t) % (uint64_t)k;
        or_overlay_run(start, len, total_len, pos0, n, out);
    }
}

#define OR_READ_TAG 0xA11CE00000000000ull
#define OR_ERR_TAG 0x5A5A5A5A00000000ull

OR_API void or_gen_reads(uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t err_den,
                         uint64_t read0, uint64_t nreads, char* out) {
    uint64_t span = genome_len - read_len + 1;
    for (uint64_t t = read0; t < read0 + nreads; t++) {
        uint64_t start = or_sm64(seed ^ (OR_READ_TAG + t)) % span;
        char* o = out + (t - read0) * ((uint64_t)read_len + 1);
        for (uint32_t j = 0; j < read_len; j++) {
            unsigned c = (unsigned)(or_sm64(seed + start + j) >> 62);
            if (err_den) {
                uint64_t e = or_sm64(seed ^ OR_ERR_TAG ^ (t * (uint64_t)read_len + j));
                if (e % err_den == 0) c = (c + 1 + (unsigned)((e >> 32) % 3)) & 3;
            }
            o[j] = "ACGT"[c];
        }
        o[read_len] = '\n';
    }
}

/* FNV-1a-64 over a byte buffer (table checksums in the golden vectors) */
OR_API uint64_t or_fnv1a64(const unsigned char* p, uint64_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}

// api.cu — context, memory helpers, k-mer enumeration, FASTA loader, text dumps.
// Host-side pieces of the C ABI (include/kmer_b200.h); no kernels here.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include <time.h>

#include "common.cuh"

static thread_local std::string g_last_error;

void kc_trace(kc_ctx* ctx, const char* what, bool sync) {
    static const bool on = getenv("KC_TRACE") != nullptr;
    if (!on) return;
    static double last = 0;
    if (sync && ctx) cudaStreamSynchronize(ctx->stream);
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    fprintf(stderr, "kc_trace %-28s +%9.3f ms\n", what, last ? now - last : 0.0);
    last = now;
}

int kc_set_error(kc_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

cudaError_t kc_pool_alloc(void** p, size_t nbytes) {
    return cudaMallocAsync(p, nbytes ? nbytes : 1, (cudaStream_t)0);
}
void kc_pool_free(void* p) {
    if (p && cudaFreeAsync(p, (cudaStream_t)0) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
    }
}
size_t kc_pool_idle_bytes(int device) {
    cudaMemPool_t pool;
    uint64_t reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess ||
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess ||
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return reserved > used ? (size_t)(reserved - used) : 0;
}

static int reserve(kc_ctx* ctx, void** p, size_t* have, size_t nbytes) {
    if (*have >= nbytes) return KC_OK;
    DeviceGuard dg(ctx->device);
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    // grow geometrically to keep reallocations rare, but never past what is needed by >25 %
    size_t want = nbytes + nbytes / 4;
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = nbytes;
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return kc_set_error(ctx, KC_ERR_NOMEM, "device scratch allocation of %zu bytes failed: %s", nbytes,
                            cudaGetErrorString(e));
    }
    *have = want;
    return KC_OK;
}
int kc_scratch_reserve(kc_ctx* ctx, size_t nbytes) { return reserve(ctx, &ctx->scratch, &ctx->scratch_bytes, nbytes); }
int kc_scratch2_reserve(kc_ctx* ctx, size_t nbytes) { return reserve(ctx, &ctx->scratch2, &ctx->scratch2_bytes, nbytes); }
void kc_scratch_release(kc_ctx* ctx) {
    DeviceGuard dg(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    ctx->scratch = ctx->scratch2 = nullptr;
    ctx->scratch_bytes = ctx->scratch2_bytes = 0;
}

extern "C" {

int kc_version(void) { return 100; }

int kc_ctx_create(int device, kc_ctx** out) {
    if (!out) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return kc_set_error(nullptr, KC_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                            e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return kc_set_error(nullptr, KC_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return kc_set_error(nullptr, KC_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return kc_set_error(nullptr, KC_ERR_UNSUPPORTED,
                            "device %d is sm_%d%d; libkmerb200 ships sm_100a code only (no fallback arch)", device,
                            prop.major, prop.minor);
    kc_ctx* ctx = new kc_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    DeviceGuard dg(device);
    // blocking streams: ordered after work the caller queued on the legacy default
    // stream (what a torch caller uses), but independent of each other
    if ((e = cudaStreamCreate(&ctx->stream)) != cudaSuccess ||
        (e = cudaStreamCreate(&ctx->copy_stream)) != cudaSuccess) {
        delete ctx;
        return kc_set_error(nullptr, KC_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    {  // freed pool memory stays with the pool (see kc_pool_alloc)
        cudaMemPool_t pool;
        uint64_t keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess ||
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
            cudaGetLastError();
    }
    *out = ctx;
    return KC_OK;
}

void kc_ctx_destroy(kc_ctx* ctx) {
    if (!ctx) return;

    DeviceGuard dg(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    {
        cudaMemPool_t pool;
        if (cudaDeviceSynchronize() != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, ctx->device) != cudaSuccess ||
            cudaMemPoolTrimTo(pool, 0) != cudaSuccess)
            cudaGetLastError();
    }
    for (auto& ev : ctx->tev)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char* kc_last_error(const kc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }
int kc_ctx_device(const kc_ctx* ctx) { return ctx ? ctx->device : -1; }
int kc_ctx_sm_count(const kc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t kc_ctx_launch_count(const kc_ctx* ctx) { return ctx ? ctx->launches : 0; }
void kc_ctx_release_memory(kc_ctx* ctx) {
    if (!ctx) return;
    kc_scratch_release(ctx);
    DeviceGuard dg(ctx->device);
    cudaMemPool_t pool;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, ctx->device) != cudaSuccess ||
        cudaMemPoolTrimTo(pool, 0) != cudaSuccess)
        cudaGetLastError();
}
void kc_ctx_set_reusable_bytes(kc_ctx* ctx, uint64_t nbytes) {
    if (ctx) ctx->caller_reusable_bytes = nbytes;
}
uint64_t kc_ctx_last_h2d_bytes(const kc_ctx* ctx) { return ctx ? ctx->last_h2d_bytes : 0; }

int kc_ctx_synchronize(kc_ctx* ctx) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return KC_OK;
}

int kc_ctx_set_timing(kc_ctx* ctx, int enabled) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    if (enabled)
        for (auto& ev : ctx->tev)
            if (!ev) KC_CUDA(ctx, cudaEventCreate(&ev));
    ctx->timing = enabled != 0;
    ctx->timed_kernels = 0;
    return KC_OK;
}

int kc_ctx_pass_times(kc_ctx* ctx, float* first_ms, float* second_ms) {
    if (!ctx || !first_ms || !second_ms) return KC_ERR_INVALID;
    *first_ms = *second_ms = 0.f;
    if (!ctx->timing || ctx->timed_kernels == 0) return KC_OK;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaEventSynchronize(ctx->tev[ctx->timed_kernels]));
    KC_CUDA(ctx, cudaEventElapsedTime(first_ms, ctx->tev[0], ctx->tev[1]));
    if (ctx->timed_kernels == 2) KC_CUDA(ctx, cudaEventElapsedTime(second_ms, ctx->tev[1], ctx->tev[2]));
    return KC_OK;
}

int kc_device_alloc(kc_ctx* ctx, size_t nbytes, void** d_out) {
    if (!ctx || !d_out) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    cudaError_t e = cudaMalloc(d_out, nbytes ? nbytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *d_out = nullptr;
        return kc_set_error(ctx, KC_ERR_NOMEM, "cudaMalloc(%zu): %s", nbytes, cudaGetErrorString(e));
    }
    return KC_OK;
}
int kc_device_free(kc_ctx* ctx, void* d_ptr) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaFree(d_ptr));
    return KC_OK;
}
int kc_host_alloc_pinned(kc_ctx* ctx, size_t nbytes, void** h_out) {
    if (!ctx || !h_out) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    cudaError_t e = cudaHostAlloc(h_out, nbytes ? nbytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *h_out = nullptr;
        return kc_set_error(ctx, KC_ERR_NOMEM, "cudaHostAlloc(%zu): %s", nbytes, cudaGetErrorString(e));
    }
    return KC_OK;
}
int kc_host_free_pinned(kc_ctx* ctx, void* h_ptr) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaFreeHost(h_ptr));
    return KC_OK;
}
int kc_memcpy_h2d(kc_ctx* ctx, void* d_dst, const void* h_src, size_t nbytes) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}
int kc_memcpy_d2h(kc_ctx* ctx, void* h_dst, const void* d_src, size_t nbytes) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}
int kc_memset_d(kc_ctx* ctx, void* d_dst, int byte, size_t nbytes) {
    if (!ctx) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaMemsetAsync(d_dst, byte, nbytes, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

// ---------------------------------------------------------------------------
// k selection / enumeration (utils.h:21-50, main.cu:122-135)
// ---------------------------------------------------------------------------
uint64_t kc_num_kmers(int k) { return (k < 1 || k > KC_MAX_K) ? 0 : (1ull << (2 * k)); }

static inline int base_code(unsigned char c) {
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

int kc_permutation(const char* alphabet, int k, char** permutations) {
    if (!alphabet || !permutations || k < 1) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_permutation: bad argument");
    const size_t na = strlen(alphabet);
    if (na < 1) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_permutation: empty alphabet");
    // |alphabet|^k entries; entry i spells the base-|alphabet| digits of i with
    // string position 0 as the LEAST significant digit (the reference's odometer
    // increments letterIdx[0] first, utils.h:36-46).
    uint64_t total = 1;
    for (int i = 0; i < k; i++) {
        if (total > (1ull << 40) / na) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_permutation: %zu^%d too large", na, k);
        total *= na;
    }
    std::vector<uint32_t> digit(k, 0);
    for (uint64_t i = 0; i < total; i++) {
        char* dst = permutations[i];
        for (int p = 0; p < k; p++) dst[p] = alphabet[digit[p]];
        dst[k] = '\0';
        for (int p = 0; p < k; p++) {
            if (++digit[p] < na) break;
            digit[p] = 0;
        }
    }
    return KC_OK;
}

int kc_kmer_index(const char* kmer, int k, uint64_t* idx_out) {
    if (!kmer || !idx_out || k < 1 || k > KC_MAX_K) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_kmer_index: bad argument");
    uint64_t v = 0;
    for (int p = 0; p < k; p++) {
        const int c = base_code((unsigned char)kmer[p]);
        if (c < 0) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_kmer_index: byte 0x%02x at %d is not in ACGT", (unsigned char)kmer[p], p);
        v |= (uint64_t)c << (2 * p);
    }
    *idx_out = v;
    return KC_OK;
}

int kc_kmer_string(uint64_t idx, int k, char* out) {
    if (!out || k < 1 || k > KC_MAX_K || (k < 32 && (idx >> (2 * k)) != 0))
        return kc_set_error(nullptr, KC_ERR_INVALID, "kc_kmer_string: bad argument");
    for (int p = 0; p < k; p++) out[p] = "ACGT"[(idx >> (2 * p)) & 3];
    out[k] = '\0';
    return KC_OK;
}

uint64_t kc_mix64(uint64_t code) { return kc_mix64_hd(code); }

int64_t kc_triangular_index(int64_t i, int64_t j, int64_t n) {
    // kernels.h:46-48 — i is 1-based, j is the gap to the later sequence
    return (n * (i - 1) - (((i - 2) * (i - 1)) / 2)) + (j - i);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// FASTA loader (importSeqs main.cu:474-545, importSeqsNoNL main.cu:401-473)
// ---------------------------------------------------------------------------
namespace {
// The output image is written once, front to back: first-touch page faults are a large part
// of the loader's time (4 KiB pages: 0.27 s per 400 MB here, and they do not scale with
// threads), so big images are 2 MiB aligned and marked for transparent huge pages.
char* alloc_image(size_t nbytes) {
    if (nbytes < (8u << 20)) return (char*)malloc(nbytes);
    const size_t huge = 2u << 20;
    const size_t rounded = (nbytes + huge - 1) / huge * huge;
    char* p = (char*)aligned_alloc(huge, rounded);
    if (p) madvise(p, rounded, MADV_HUGEPAGE);
    return p;
}

// std::getline over a memory image: yields every '\n'-terminated line plus a
// final unterminated one, like the ifstream loop at main.cu:487.
struct LineReader {
    const char* p;
    const char* end;
    bool next(const char*& line, size_t& len) {
        if (p >= end) return false;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        line = p;
        if (nl) {
            len = (size_t)(nl - p);
            p = nl + 1;
        } else {
            len = (size_t)(end - p);
            p = end;
        }
        return true;
    }
};

// A record is built in place at the tail of s->data (no per-record temporary):
// `rec_begin` is where the open record starts.  data has room for the whole file
// plus one separator per record (a record needs at least a header line, i.e. two
// bytes of input that are not copied, for the one byte it adds).
inline void put(kc_seqset* s, const char* p, size_t n) {
    memcpy(s->data + s->data_len, p, n);
    s->data_len += n;
}

void close_record(kc_seqset* s, size_t rec_begin) {
    s->offsets.push_back((int64_t)rec_begin);
    // main.cu:538-541 turns every '|' of globalAcc into NUL
    char* p = s->data + rec_begin;
    char* const end = s->data + s->data_len;
    while ((p = (char*)memchr(p, '|', (size_t)(end - p))) != nullptr) *p++ = '\0';
    s->data[s->data_len++] = '\0';  // the record's own '|' separator (main.cu:505,517)
    s->num_seqs++;
}

int parse_fasta(const char* img, size_t n, int mode, long max_seqs, kc_seqset* s) {
    LineReader rd{img, img + n};
    const char* line;
    size_t len;
    bool armed = false, stop = false;
    s->data_cap = n + 16;
    s->data = alloc_image(s->data_cap);
    if (!s->data) return kc_set_error(nullptr, KC_ERR_NOMEM, "kc_import_seqs: cannot allocate %zu bytes", s->data_cap);
    while (!stop && rd.next(line, len)) {
        if (len == 0) continue;  // blank line between records
        if (line[0] == '>') {
            s->ids.emplace_back(line, len);
            armed = true;
            continue;
        }
        if (!armed) continue;  // sequence text without a header is dropped
        armed = false;
        const size_t rec_begin = s->data_len;
        put(s, line, len);
        bool closed = false;
        while (rd.next(line, len)) {
            const bool header = (mode == KC_IMPORT_NONL && len > 0 && line[0] == '>');
            if (header) armed = true;  // NoNL: the header ends the record but is not kept
            if (len == 0 || line[0] == '\r' || header) {
                close_record(s, rec_begin);
                closed = true;
                break;
            }
            put(s, line, len);
            if (max_seqs > 0 && (long)s->num_seqs >= max_seqs) break;
        }
        if (!closed && s->data_len > rec_begin) {
            close_record(s, rec_begin);
            if (max_seqs > 0 && (long)s->num_seqs >= max_seqs) stop = true;
        }
    }
    s->offsets.push_back((int64_t)s->data_len);  // terminal offset, always
    return KC_OK;
}
// ---------------------------------------------------------------------------
// The same loader on several host threads ("next" row f2: ingest at speed).
// The line loop above is a three-state machine — IDLE (no header seen / record closed),
// ARMED (header seen, waiting for the first sequence line), INREC (inside a record) — whose
// transitions depend only on the class of each line (empty, '>', '\r', other) and the mode.
// So the image is cut at line starts into one chunk per thread and
//   pass 1: every thread runs its chunk from ALL THREE start states at once and records, per
//           start state: end state, bytes emitted, records started, ids pushed;
//   stitch: the true start state / output offset / record index of every chunk follow serially;
//   pass 2: every thread runs its chunk again from its true start state and writes
//           data, offsets and ids at its own positions.
// Output is byte-identical to parse_fasta with max_seqs <= 0 (tests/test_host_api.py compares
// them on the reference fixtures and on random files with 1..9 threads).
// ---------------------------------------------------------------------------
enum { ST_IDLE = 0, ST_ARMED = 1, ST_INREC = 2 };

struct ThreadGroup {  // joins whatever was started, also when starting the next thread throws
    std::vector<std::thread> th;
    ~ThreadGroup() {
        for (auto& t : th)
            if (t.joinable()) t.join();
    }
};

struct ChunkSummary {
    int end_state[3];
    uint64_t bytes[3], recs[3], ids[3];
};

// One line through the machine.  Returns the new state; *copy = the line's bytes are
// sequence data; *close = a record ends BEFORE this line (one separator byte); *start = a
// record starts with this line; *id = the line is pushed to ids.
inline int step_line(int st, const char* line, size_t len, int mode, bool* copy, bool* close, bool* start, bool* id) {
    *copy = *close = *start = *id = false;
    const bool empty = len == 0, header = !empty && line[0] == '>', cr = !empty && line[0] == '\r';
    if (st == ST_INREC) {
        if (empty || cr) {
            *close = true;
            return ST_IDLE;
        }
        if (header && mode == KC_IMPORT_NONL) {  // ends the record, arms the next one, is not kept
            *close = true;
            return ST_ARMED;
        }
        *copy = true;  // text, or (mode 0) a '>' line inside a record: appended
        return ST_INREC;
    }
    if (empty) return st;
    if (header) {
        *id = true;
        return ST_ARMED;
    }
    if (st == ST_IDLE) return ST_IDLE;  // sequence text without a header is dropped
    *copy = *start = true;              // ARMED: first line of a record (even a '\r' line)
    return ST_INREC;
}

void summarize_chunk(const char* b, const char* e, int mode, ChunkSummary* out) {
    int st[3] = {ST_IDLE, ST_ARMED, ST_INREC};
    uint64_t bytes[3] = {0, 0, 0}, recs[3] = {0, 0, 0}, ids[3] = {0, 0, 0};
    LineReader rd{b, e};
    const char* line;
    size_t len;
    while (rd.next(line, len)) {
        for (int s = 0; s < 3; s++) {
            bool copy, close, start, id;
            st[s] = step_line(st[s], line, len, mode, &copy, &close, &start, &id);
            bytes[s] += (copy ? len : 0) + (close ? 1 : 0);
            recs[s] += start ? 1 : 0;
            ids[s] += id ? 1 : 0;
        }
    }
    for (int s = 0; s < 3; s++) {
        out->end_state[s] = st[s];
        out->bytes[s] = bytes[s];
        out->recs[s] = recs[s];
        out->ids[s] = ids[s];
    }
}

void emit_chunk(const char* b, const char* e, int mode, int st, char* data, uint64_t pos, int64_t* offsets, uint64_t rec,
                std::string* ids) {
    LineReader rd{b, e};
    const char* line;
    size_t len;
    while (rd.next(line, len)) {
        bool copy, close, start, id;
        st = step_line(st, line, len, mode, &copy, &close, &start, &id);
        if (close) data[pos++] = '\0';  // the record's own '|' separator (main.cu:505,517), already NUL
        if (start) offsets[rec++] = (int64_t)pos;
        if (copy) {
            memcpy(data + pos, line, len);
            char* p = data + pos;  // main.cu:538-541 turns every '|' into NUL
            char* const end = p + len;
            while ((p = (char*)memchr(p, '|', (size_t)(end - p))) != nullptr) *p++ = '\0';
            pos += len;
        }
        if (id) (ids++)->assign(line, len);
    }
}

int parse_fasta_threads(const char* img, size_t n, int mode, int nthreads, kc_seqset* s) {
    if (nthreads < 1) nthreads = 1;
    // chunk starts at line starts
    std::vector<size_t> cut(1, 0);
    for (int t = 1; t < nthreads; t++) {
        size_t at = n / nthreads * t;
        if (at <= cut.back()) continue;
        const char* nl = (const char*)memchr(img + at, '\n', n - at);
        if (!nl) break;
        at = (size_t)(nl - img) + 1;
        if (at > cut.back() && at < n) cut.push_back(at);
    }
    cut.push_back(n);
    const int nc = (int)cut.size() - 1;
    std::vector<ChunkSummary> sum(nc);
    const bool dbg = getenv("KC_LOADER_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_a = now();
    {
        ThreadGroup g;
        for (int c = 1; c < nc; c++) g.th.emplace_back(summarize_chunk, img + cut[c], img + cut[c + 1], mode, &sum[c]);
        summarize_chunk(img + cut[0], img + cut[1], mode, &sum[0]);
    }
    const double t_b = now();
    std::vector<int> st0(nc);
    std::vector<uint64_t> pos0(nc), rec0(nc), id0(nc);
    int st = ST_IDLE;
    uint64_t pos = 0, rec = 0, nid = 0;
    for (int c = 0; c < nc; c++) {
        st0[c] = st;
        pos0[c] = pos;
        rec0[c] = rec;
        id0[c] = nid;
        pos += sum[c].bytes[st];
        rec += sum[c].recs[st];
        nid += sum[c].ids[st];
        st = sum[c].end_state[st];
    }
    const bool open_at_eof = (st == ST_INREC);  // the last record is closed by EOF (main.cu:516-524)
    const uint64_t total = pos + (open_at_eof ? 1 : 0);
    if (rec > 0xFFFFFFFFull) return kc_set_error(nullptr, KC_ERR_UNSUPPORTED, "kc_import_seqs: more than 2^32-1 records");
    s->data_cap = total + 16;
    s->data = alloc_image(s->data_cap);
    if (!s->data) return kc_set_error(nullptr, KC_ERR_NOMEM, "kc_import_seqs: cannot allocate %zu bytes", s->data_cap);
    s->offsets.assign(rec + 1, 0);
    s->ids.assign(nid, std::string());
    const double t_c = now();
    std::atomic<int> emit_failed{0};
    auto emit_guarded = [&](int c) {
        try {
            emit_chunk(img + cut[c], img + cut[c + 1], mode, st0[c], s->data, pos0[c], s->offsets.data(), rec0[c], s->ids.data() + id0[c]);
        } catch (...) {  // std::string growth of an id: the only allocation in pass 2
            emit_failed = 1;
        }
    };
    {
        ThreadGroup g;
        for (int c = 1; c < nc; c++) g.th.emplace_back(emit_guarded, c);
        emit_guarded(0);
    }
    if (emit_failed) return kc_set_error(nullptr, KC_ERR_NOMEM, "kc_import_seqs: out of memory while copying record ids");
    if (dbg)
        fprintf(stderr, "kc_import_seqs: %d chunks, pass 1 %.3f s, stitch+alloc %.3f s, pass 2 %.3f s\n", nc, t_b - t_a, t_c - t_b,
                now() - t_c);
    if (open_at_eof) s->data[pos] = '\0';
    s->data_len = total;
    s->num_seqs = (uint32_t)rec;
    s->offsets[rec] = (int64_t)total;  // terminal offset, always
    return KC_OK;
}
}  // namespace

extern "C" {

int kc_import_seqs_mem_threads(const char* fasta, size_t nbytes, int mode, int nthreads, kc_seqset** out) {
    if (!out || (!fasta && nbytes)) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_import_seqs_mem_threads: null pointer");
    *out = nullptr;
    if (mode != KC_IMPORT_BLANKLINE && mode != KC_IMPORT_NONL)
        return kc_set_error(nullptr, KC_ERR_INVALID, "kc_import_seqs: unknown mode %d", mode);
    if (nthreads <= 0) {
        nthreads = (int)std::thread::hardware_concurrency();
        if (nthreads < 1) nthreads = 1;
        if (nthreads > 64) nthreads = 64;
        const size_t by_size = nbytes / (4u << 20) + 1;  // at least 4 MiB of text per thread
        if ((size_t)nthreads > by_size) nthreads = (int)by_size;
    }
    kc_seqset* s = new kc_seqset();
    int rc;
    try {  // no C++ exception may cross the C ABI (thread creation, allocation of ids / offsets)
        rc = parse_fasta_threads(fasta, nbytes, mode, nthreads, s);
    } catch (const std::exception& e) {
        rc = kc_set_error(nullptr, KC_ERR_NOMEM, "kc_import_seqs: %s", e.what());
    }
    if (rc) {
        delete s;
        return rc;
    }
    *out = s;
    return KC_OK;
}

int kc_import_seqs_mem(const char* fasta, size_t nbytes, int mode, long max_seqs, kc_seqset** out) {
    // no record limit and enough text: the multi-threaded parser (same result)
    if (max_seqs <= 0 && nbytes >= (32u << 20)) return kc_import_seqs_mem_threads(fasta, nbytes, mode, 0, out);
    if (!out || (!fasta && nbytes)) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_import_seqs_mem: null pointer");
    if (mode != KC_IMPORT_BLANKLINE && mode != KC_IMPORT_NONL)
        return kc_set_error(nullptr, KC_ERR_INVALID, "kc_import_seqs: unknown mode %d", mode);
    kc_seqset* s = new kc_seqset();
    int rc;
    try {
        rc = parse_fasta(fasta, nbytes, mode, max_seqs, s);
    } catch (const std::exception& e) {
        rc = kc_set_error(nullptr, KC_ERR_NOMEM, "kc_import_seqs: %s", e.what());
    }
    if (rc) {
        delete s;
        return rc;
    }
    *out = s;
    return KC_OK;
}

int kc_import_seqs(const char* path, int mode, long max_seqs, kc_seqset** out) {
    if (!path || !out) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_import_seqs: null pointer");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    // the reference prints "Error opening" and exit(0)s here (main.cu:477-480)
    if (!f) return kc_set_error(nullptr, KC_ERR_IO, "Error opening: %s . Check your file or path.", path);
    // map the file when it has a size (no copy of the input); otherwise read it
    struct stat stt;
    const int fd = fileno(f);
    if (fstat(fd, &stt) == 0 && S_ISREG(stt.st_mode) && stt.st_size > 0) {
        void* m = mmap(nullptr, (size_t)stt.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)stt.st_size, MADV_SEQUENTIAL);
            const int rc = kc_import_seqs_mem((const char*)m, (size_t)stt.st_size, mode, max_seqs, out);
            munmap(m, (size_t)stt.st_size);
            fclose(f);
            return rc;
        }
    }
    std::string img;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) img.append(buf, got);
    fclose(f);
    return kc_import_seqs_mem(img.data(), img.size(), mode, max_seqs, out);
}

void kc_seqset_free(kc_seqset* s) {
    if (!s) return;
    if (s->device >= 0) {  // the ctx that made the device copies may already be destroyed: only its device index is used
        DeviceGuard dg(s->device);
        if (s->d_data) cudaFree(s->d_data);
        if (s->d_offsets) cudaFree(s->d_offsets);
    }
    delete s;
}
uint32_t kc_seqset_num_seqs(const kc_seqset* s) { return s ? s->num_seqs : 0; }
uint32_t kc_seqset_num_ids(const kc_seqset* s) { return s ? (uint32_t)s->ids.size() : 0; }
uint64_t kc_seqset_nbytes(const kc_seqset* s) { return s ? s->data_len : 0; }
const char* kc_seqset_data(const kc_seqset* s) {
    if (!s) return nullptr;
    // a set parsed on the GPU (kc_import_seqs_device) fetches its host image on first use
    if (!s->data && s->d_data && s->device >= 0 && s->data_len) {
        kc_seqset* m = const_cast<kc_seqset*>(s);
        DeviceGuard dg(m->device);
        char* h = (char*)malloc(m->data_len + 16);
        if (!h) return nullptr;
        if (cudaMemcpy(h, m->d_data, m->data_len, cudaMemcpyDeviceToHost) != cudaSuccess) {  // synchronous: no ctx stream needed
            cudaGetLastError();
            free(h);
            return nullptr;
        }
        m->data = h;
        m->data_cap = m->data_len + 16;
    }
    return s->data;
}
const int64_t* kc_seqset_offsets(const kc_seqset* s) { return s ? s->offsets.data() : nullptr; }
const char* kc_seqset_id(const kc_seqset* s, uint32_t i) {
    return (s && i < s->ids.size()) ? s->ids[i].c_str() : nullptr;
}

int kc_seqset_to_device(kc_ctx* ctx, kc_seqset* s, const char** d_data, const int64_t** d_offsets) {
    if (!ctx || !s) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    if (!s->d_data) {
        s->owner = ctx;
        s->device = ctx->device;
        KC_CUDA(ctx, cudaMalloc(&s->d_data, s->data_len ? s->data_len : 1));
        KC_CUDA(ctx, cudaMalloc(&s->d_offsets, s->offsets.size() * sizeof(int64_t)));
        KC_CUDA(ctx, cudaMemcpyAsync(s->d_data, s->data, s->data_len, cudaMemcpyHostToDevice, ctx->stream));
        KC_CUDA(ctx, cudaMemcpyAsync(s->d_offsets, s->offsets.data(), s->offsets.size() * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, ctx->stream));
        KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (d_data) *d_data = s->d_data;
    if (d_offsets) *d_offsets = s->d_offsets;
    return KC_OK;
}

// ---------------------------------------------------------------------------
// text dumps (main.cu:301-309 and main.cu:355-358)
// ---------------------------------------------------------------------------
int kc_dump_counts(const char* path, const int32_t* h_sums, int k, uint32_t num_seqs) {
    if (!h_sums || k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_dump_counts: bad argument");
    FILE* f = path ? fopen(path, "w") : stdout;
    if (!f) return kc_set_error(nullptr, KC_ERR_IO, "kc_dump_counts: cannot open %s", path);
    const uint64_t nk = 1ull << (2 * k);
    fprintf(f, "Sums:\n");
    uint64_t idx = 0;
    for (uint64_t j = 0; j < nk; j++) {
        fprintf(f, "%d: ", (int)j);
        for (uint32_t i = 0; i < num_seqs; i++) fprintf(f, "%d,\t", h_sums[idx++]);
        fprintf(f, "\n");
    }
    fprintf(f, "\n");
    if (path)
        fclose(f);
    else
        fflush(f);
    return KC_OK;
}

int kc_dump_distances(const char* path, const float* h_dist, uint64_t n_pairs) {
    if (!h_dist && n_pairs) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_dump_distances: null pointer");
    FILE* f = path ? fopen(path, "w") : stdout;
    if (!f) return kc_set_error(nullptr, KC_ERR_IO, "kc_dump_distances: cannot open %s", path);
    for (uint64_t i = 0; i < n_pairs; i++) fprintf(f, "%f\n", h_dist[i]);
    if (path)
        fclose(f);
    else
        fflush(f);
    return KC_OK;
}

}  // extern "C"

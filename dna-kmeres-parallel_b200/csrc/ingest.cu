// ingest.cu — "next" row f2, device side: importSeqs / importSeqsNoNL (main.cu:474-545, 401-473) on the
// GPU.  The FASTA file image goes to HBM as it is on disk; newlines, header lines and dropped lines
// are stripped THERE, so the host never touches the sequence bytes (first B200 run pending).
//
// The loader is a finite-state transducer over BYTES with seven states:
//     0  line start, no record open ("idle")        3  inside a line that is dropped, then idle
//     1  line start, header seen ("armed")          4  inside a header line (an id), then armed
//     2  line start, record open                    5  inside a line that is copied, then record open
//                                                   6  inside a dropped header (NoNL), then armed
// (transitions in fsm_step: the rules of SURVEY §8a / kc_import_seqs, byte by byte).  A transducer's
// effect on a tile of bytes is a map  start state -> (end state, bytes emitted, records started,
// ids seen), and maps compose associatively, so:
//   K_A  one thread per 8 KiB tile: the tile's map.  Seven machines run only up to the tile's first
//        newline — there every one of them is in state 0, 1 or 2 — and three canonical machines
//        do the rest.
//   K_B  one CTA composes the maps (each thread a run of tiles, thread 0 the 1024 runs, each thread
//        its run again): start state, output offset, record and id index of every tile.
//   K_C  one thread per tile runs the transducer from its true start state and writes the sequence
//        bytes ('|' -> NUL, main.cu:538-541), the separators, the record offsets and the file
//        positions of the id lines (the id strings are cut from the host's file image).
// Output identical to kc_import_seqs_mem with max_seqs <= 0 (tests/test_emu_kernels.py compares them
// on the reference fixtures and random files with tiles of 16 bytes).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.cuh"

namespace {

constexpr int NSTATE = 7;

struct Effect {
    uint32_t copy, close, start, id;  // this byte is sequence data / a record ends before it / starts with it / an id starts
};

__host__ __device__ inline int fsm_step(int st, unsigned char c, int mode, Effect& e) {
    e.copy = e.close = e.start = e.id = 0;
    const bool nl = (c == '\n');
    switch (st) {
        case 0:  // line start, idle
            if (nl) return 0;
            if (c == '>') {
                e.id = 1;
                return 4;
            }
            return 3;  // sequence text without a header is dropped
        case 1:  // line start, armed
            if (nl) return 1;  // blank lines between header and sequence are skipped
            if (c == '>') {
                e.id = 1;
                return 4;
            }
            e.start = e.copy = 1;  // first line of a record (even a '\r' line)
            return 5;
        case 2:  // line start, record open
            if (nl) {
                e.close = 1;  // a blank line ends the record
                return 0;
            }
            if (c == '\r') {
                e.close = 1;  // so does a line that starts with CR; its rest is dropped
                return 3;
            }
            if (c == '>' && mode == KC_IMPORT_NONL) {
                e.close = 1;  // NoNL: a header ends the record, arms the next one, is not kept
                return 6;
            }
            e.copy = 1;  // text, or (mode 0) a '>' line inside a record: appended
            return 5;
        case 3:
            return nl ? 0 : 3;
        case 4:
            return nl ? 1 : 4;
        case 5:
            if (nl) return 2;
            e.copy = 1;
            return 5;
        default:  // 6
            return nl ? 1 : 6;
    }
}

// bytes [b, e) of raw, in order; 128-bit loads wherever a 16-byte aligned block lies inside the range
// (a thread walks its own tile: byte loads would cost one LSU wavefront per byte and lane)
template <typename F>
__device__ __forceinline__ void for_each_byte(const unsigned char* __restrict__ raw, uint64_t b, uint64_t e, F&& f) {
    uint64_t p = b;
    while (p < e) {
        if ((((uintptr_t)(raw + p)) & 15) == 0 && p + 16 <= e) {
            const uint4 v = *reinterpret_cast<const uint4*>(raw + p);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; i++) f((unsigned char)(w[i >> 2] >> (8 * (i & 3))), p + i);
            p += 16;
        } else {
            f(raw[p], p);
            p++;
        }
    }
}

struct TileMap {  // start state -> effect of the whole tile
    uint32_t bytes[NSTATE], recs[NSTATE], ids[NSTATE];
    uint8_t end[8];
};

__global__ void __launch_bounds__(256)
ingest_map_kernel(const unsigned char* __restrict__ raw, uint64_t n, uint32_t tile, uint64_t ntiles, int mode, TileMap* __restrict__ maps) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = t * tile, e = (b + tile < n) ? b + tile : n;
        int st7[NSTATE];
        uint32_t by7[NSTATE], rc7[NSTATE], id7[NSTATE];
#pragma unroll
        for (int s = 0; s < NSTATE; s++) {
            st7[s] = s;
            by7[s] = rc7[s] = id7[s] = 0;
        }
        bool merged = false;  // past the first newline: every machine is in state 0, 1 or 2
        int stc[3] = {0, 1, 2};
        uint32_t byc[3] = {0, 0, 0}, rcc[3] = {0, 0, 0}, idc[3] = {0, 0, 0};
        auto feed = [&](unsigned char c, uint64_t) {
            if (!merged) {
#pragma unroll
                for (int s = 0; s < NSTATE; s++) {
                    Effect ef;
                    st7[s] = fsm_step(st7[s], c, mode, ef);
                    by7[s] += ef.copy + ef.close;
                    rc7[s] += ef.start;
                    id7[s] += ef.id;
                }
                merged = (c == '\n');
            } else {
#pragma unroll
                for (int s = 0; s < 3; s++) {
                    Effect ef;
                    stc[s] = fsm_step(stc[s], c, mode, ef);
                    byc[s] += ef.copy + ef.close;
                    rcc[s] += ef.start;
                    idc[s] += ef.id;
                }
            }
        };
        for_each_byte(raw, b, e, feed);
        TileMap m;
#pragma unroll
        for (int s = 0; s < NSTATE; s++) {
            int mid = st7[s];
            if (merged) {  // mid is 0, 1 or 2: continue with that canonical machine (static indexing: registers)
                const int es = mid == 0 ? stc[0] : mid == 1 ? stc[1] : stc[2];
                m.end[s] = (uint8_t)es;
                m.bytes[s] = by7[s] + (mid == 0 ? byc[0] : mid == 1 ? byc[1] : byc[2]);
                m.recs[s] = rc7[s] + (mid == 0 ? rcc[0] : mid == 1 ? rcc[1] : rcc[2]);
                m.ids[s] = id7[s] + (mid == 0 ? idc[0] : mid == 1 ? idc[1] : idc[2]);
            } else {
                m.end[s] = (uint8_t)mid;
                m.bytes[s] = by7[s];
                m.recs[s] = rc7[s];
                m.ids[s] = id7[s];
            }
        }
        m.end[7] = 0;
        maps[t] = m;
    }
}

struct TileStart {
    uint64_t out_pos, rec, id;
    uint32_t state, pad;
};
struct IngestTotals {
    uint64_t bytes, recs, ids;
    uint32_t end_state, pad;
};

struct RunMap {  // composition of a run of tiles
    uint64_t bytes[NSTATE], recs[NSTATE], ids[NSTATE];
    int end[NSTATE];
};

// one CTA of 1024 threads; thread r owns tiles [r * per, (r+1) * per)
__global__ void __launch_bounds__(1024, 1)
ingest_scan_kernel(const TileMap* __restrict__ maps, uint64_t ntiles, TileStart* __restrict__ starts, IngestTotals* totals,
                   RunMap* __restrict__ runs /* [1024] scratch */) {
    const int r = threadIdx.x;
    const uint64_t per = (ntiles + 1023) / 1024;
    const uint64_t b = min((uint64_t)r * per, ntiles), e = min(b + per, ntiles);
    {
        RunMap m;
        for (int s = 0; s < NSTATE; s++) {
            m.end[s] = s;
            m.bytes[s] = m.recs[s] = m.ids[s] = 0;
        }
        for (uint64_t t = b; t < e; t++) {
            const TileMap& tm = maps[t];
            for (int s = 0; s < NSTATE; s++) {
                const int mid = m.end[s];
                m.bytes[s] += tm.bytes[mid];
                m.recs[s] += tm.recs[mid];
                m.ids[s] += tm.ids[mid];
                m.end[s] = tm.end[mid];
            }
        }
        runs[r] = m;
    }
    __syncthreads();
    __shared__ uint64_t s_bytes[1024], s_recs[1024], s_ids[1024];
    __shared__ int s_state[1024];
    if (r == 0) {
        int st = 0;  // the file starts at a line start with no record open
        uint64_t by = 0, rc = 0, id = 0;
        for (int q = 0; q < 1024; q++) {
            s_state[q] = st;
            s_bytes[q] = by;
            s_recs[q] = rc;
            s_ids[q] = id;
            const RunMap& m = runs[q];
            by += m.bytes[st];
            rc += m.recs[st];
            id += m.ids[st];
            st = m.end[st];
        }
        totals->bytes = by;
        totals->recs = rc;
        totals->ids = id;
        totals->end_state = (uint32_t)st;
    }
    __syncthreads();
    int st = s_state[r];
    uint64_t by = s_bytes[r], rc = s_recs[r], id = s_ids[r];
    for (uint64_t t = b; t < e; t++) {
        TileStart ts;
        ts.out_pos = by;
        ts.rec = rc;
        ts.id = id;
        ts.state = (uint32_t)st;
        ts.pad = 0;
        starts[t] = ts;
        const TileMap& tm = maps[t];
        by += tm.bytes[st];
        rc += tm.recs[st];
        id += tm.ids[st];
        st = tm.end[st];
    }
}

__global__ void __launch_bounds__(256)
ingest_emit_kernel(const unsigned char* __restrict__ raw, uint64_t n, uint32_t tile, uint64_t ntiles, int mode,
                   const TileStart* __restrict__ starts, char* __restrict__ data, int64_t* __restrict__ offsets,
                   uint64_t* __restrict__ id_pos) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = t * tile, e = (b + tile < n) ? b + tile : n;
        const TileStart ts = starts[t];
        int st = (int)ts.state;
        uint64_t o = ts.out_pos, rc = ts.rec, id = ts.id;
        // output bytes are gathered into aligned 32-bit words (a byte store costs a wavefront per lane)
        uint32_t word = 0, fill = 0;  // `fill` bytes of the word at o - fill are pending; o - fill is 4-byte aligned
        auto put = [&](unsigned char ch) {
            if (fill == 0 && (((uintptr_t)(data + o)) & 3)) {
                data[o++] = (char)ch;  // head bytes up to the first aligned word
                return;
            }
            word |= (uint32_t)ch << (8 * fill);
            fill++;
            o++;
            if (fill == 4) {
                *reinterpret_cast<uint32_t*>(data + o - 4) = word;
                word = 0;
                fill = 0;
            }
        };
        for_each_byte(raw, b, e, [&](unsigned char c, uint64_t p) {
            Effect ef;
            st = fsm_step(st, c, mode, ef);
            if (ef.close) put(0);  // the record's own '|' separator (main.cu:505,517), already NUL
            if (ef.start) offsets[rc++] = (int64_t)o;
            if (ef.copy) put(c == '|' ? (unsigned char)0 : c);  // main.cu:538-541
            if (ef.id) id_pos[id++] = p;
        });
        for (uint32_t q = 0; q < fill; q++) data[o - fill + q] = (char)(word >> (8 * q));  // tail bytes
        if (e == n && (st == 2 || st == 5)) data[o] = '\0';  // the last record is closed by the end of the file
    }
}

struct Dev {
    void* p = nullptr;
    ~Dev() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    void* release() {
        void* q = p;
        p = nullptr;
        return q;
    }
};

}  // namespace

extern "C" int kc_import_seqs_device(kc_ctx* ctx, const char* d_raw, const char* h_raw, uint64_t nbytes, int mode, kc_seqset** out) {
    if (!ctx || !out) return KC_ERR_INVALID;
    *out = nullptr;
    if (mode != KC_IMPORT_BLANKLINE && mode != KC_IMPORT_NONL) return kc_set_error(ctx, KC_ERR_INVALID, "kc_import_seqs: unknown mode %d", mode);
    if (!d_raw && nbytes) return kc_set_error(ctx, KC_ERR_INVALID, "kc_import_seqs_device: null pointer");
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    kc_seqset* s = nullptr;
    try {
        s = new kc_seqset();
        s->owner = ctx;
        s->device = ctx->device;
        static const uint32_t tile_env = getenv("KC_INGEST_TILE") ? (uint32_t)atoi(getenv("KC_INGEST_TILE")) : 0;  // test aid
        const uint32_t tile = tile_env ? tile_env : 8192u;
        const uint64_t ntiles = (nbytes + tile - 1) / tile;
        if (ntiles == 0) {
            s->offsets.assign(1, 0);
            *out = s;
            return KC_OK;
        }
        Dev maps, starts, totals, runs;
        if (maps.alloc(ntiles * sizeof(TileMap)) || starts.alloc(ntiles * sizeof(TileStart)) || totals.alloc(sizeof(IngestTotals)) ||
            runs.alloc(1024 * sizeof(RunMap))) {
            cudaGetLastError();
            delete s;
            return kc_set_error(ctx, KC_ERR_NOMEM, "kc_import_seqs_device: out of device memory");
        }
        const uint64_t want = (ntiles + 255) / 256;
        const int grid = (int)(want > (uint64_t)ctx->sm_count * 8 ? (uint64_t)ctx->sm_count * 8 : want);
        const unsigned char* raw = (const unsigned char*)d_raw;
        int rc = KC_OK;
        auto run = [&]() -> int {
            KC_LAUNCH(ingest_map_kernel, grid, 256, 0, st, raw, nbytes, tile, ntiles, mode, (TileMap*)maps.p);
            KC_LAUNCH_CHECK(ctx, "ingest_map_kernel");
            KC_LAUNCH(ingest_scan_kernel, 1, 1024, 0, st, (const TileMap*)maps.p, ntiles, (TileStart*)starts.p, (IngestTotals*)totals.p,
                      (RunMap*)runs.p);
            KC_LAUNCH_CHECK(ctx, "ingest_scan_kernel");
            IngestTotals h;
            KC_CUDA(ctx, cudaMemcpyAsync(&h, totals.p, sizeof h, cudaMemcpyDeviceToHost, st));
            KC_CUDA(ctx, cudaStreamSynchronize(st));
            const bool open_at_eof = (h.end_state == 2 || h.end_state == 5);
            const uint64_t total = h.bytes + (open_at_eof ? 1 : 0);
            if (h.recs > 0xFFFFFFFFull) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "kc_import_seqs: more than 2^32-1 records");
            Dev idpos;
            if (idpos.alloc(h.ids * 8)) {
                cudaGetLastError();
                return kc_set_error(ctx, KC_ERR_NOMEM, "kc_import_seqs_device: out of device memory");
            }
            KC_CUDA(ctx, cudaMalloc(&s->d_data, total ? total : 1));
            KC_CUDA(ctx, cudaMalloc(&s->d_offsets, (h.recs + 1) * sizeof(int64_t)));
            KC_LAUNCH(ingest_emit_kernel, grid, 256, 0, st, raw, nbytes, tile, ntiles, mode, (const TileStart*)starts.p, s->d_data,
                      s->d_offsets, (uint64_t*)idpos.p);
            KC_LAUNCH_CHECK(ctx, "ingest_emit_kernel");
            const int64_t term = (int64_t)total;  // terminal offset, always
            KC_CUDA(ctx, cudaMemcpyAsync(s->d_offsets + h.recs, &term, sizeof term, cudaMemcpyHostToDevice, st));
            s->offsets.assign(h.recs + 1, 0);
            KC_CUDA(ctx, cudaMemcpyAsync(s->offsets.data(), s->d_offsets, (h.recs + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            std::vector<uint64_t> pos(h.ids);
            if (h.ids) KC_CUDA(ctx, cudaMemcpyAsync(pos.data(), idpos.p, h.ids * 8, cudaMemcpyDeviceToHost, st));
            KC_CUDA(ctx, cudaStreamSynchronize(st));
            s->data_len = total;
            s->num_seqs = (uint32_t)h.recs;
            if (h_raw) {  // the id strings are the header lines of the host's copy of the file
                s->ids.reserve(h.ids);
                for (uint64_t q = 0; q < h.ids; q++) {
                    const char* b = h_raw + pos[q];
                    const char* nl = (const char*)memchr(b, '\n', (size_t)(nbytes - pos[q]));
                    s->ids.emplace_back(b, nl ? (size_t)(nl - b) : (size_t)(nbytes - pos[q]));
                }
            }
            return KC_OK;
        };
        rc = run();
        if (rc) {
            kc_seqset_free(s);
            return rc;
        }
    } catch (const std::exception& e) {
        if (s) kc_seqset_free(s);
        return kc_set_error(ctx, KC_ERR_NOMEM, "kc_import_seqs_device: %s", e.what());
    }
    *out = s;
    return KC_OK;
}

// File form: the file is mapped, copied to the device in 64 MiB pieces and parsed there; the id
// strings come from the mapping.  Error text for a missing file as in the reference (main.cu:477-480).
extern "C" int kc_import_seqs_gpu(kc_ctx* ctx, const char* path, int mode, kc_seqset** out) {
    if (!ctx || !out || !path) return KC_ERR_INVALID;
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return kc_set_error(ctx, KC_ERR_IO, "Error opening: %s . Check your file or path.", path);
    struct stat stt;
    if (fstat(fd, &stt) != 0 || !S_ISREG(stt.st_mode)) {
        close(fd);
        return kc_set_error(ctx, KC_ERR_IO, "kc_import_seqs_gpu: %s is not a regular file", path);
    }
    const uint64_t n = (uint64_t)stt.st_size;
    if (n == 0) {
        close(fd);
        return kc_import_seqs_device(ctx, nullptr, nullptr, 0, mode, out);
    }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return kc_set_error(ctx, KC_ERR_IO, "kc_import_seqs_gpu: cannot map %s", path);
    madvise(m, n, MADV_SEQUENTIAL);
    DeviceGuard dg(ctx->device);
    Dev raw;
    int rc = KC_OK;
    if (raw.alloc(n)) {
        cudaGetLastError();
        rc = kc_set_error(ctx, KC_ERR_NOMEM, "kc_import_seqs_gpu: no device memory for a %llu-byte file", (unsigned long long)n);
    }
    const uint64_t piece = 64ull << 20;
    for (uint64_t b = 0; b < n && rc == KC_OK; b += piece) {
        const uint64_t len = (b + piece < n) ? piece : n - b;
        if (cudaMemcpyAsync((char*)raw.p + b, (const char*)m + b, len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            rc = kc_set_error(ctx, KC_ERR_CUDA, "kc_import_seqs_gpu: copy to the device failed");
        }
    }
    if (rc == KC_OK) rc = kc_import_seqs_device(ctx, (const char*)raw.p, (const char*)m, n, mode, out);
    cudaStreamSynchronize(ctx->stream);
    munmap(m, n);
    return rc;
}

// sparse.cu — sparse k-mer counting for k <= 31 (uint64 little-endian codes).
//
// The reference cannot express k > 6 at all (kernels.h:21: c_perms overflows
// constant memory; the kernel hard-codes 3, kernels.h:133-138).  These kernels
// extend the same window semantics (main.cu:636-646) to 64-bit codes:
//
//   KC_SPARSE_HASH  open-addressing table in HBM: 16-byte slots {uint64 key, uint32 count}
//                   (CAS on the key, linear probing from mix64(code), RED.ADD on the count),
//                   filled straight from the WarpScanner — codes never touch
//                   memory; then non-empty slots are compacted and sorted.
//   KC_SPARSE_SORT  codes written per window chunk, radix-sorted (CUB, a
//                   library sort), run-length reduced, chunk results merged by a
//                   second sort + reduce-by-key.
//
// Result in both cases: distinct codes ascending + counts (kc_sparse).
#ifndef KC_EMU
#include <cub/cub.cuh>
#else
#include "../../tests/emu/cub_emu.h"  // test-only CPU emulator build
#endif

#include "common.cuh"

static constexpr uint64_t KEY_EMPTY = ~0ull;

// ---------------------------------------------------------------------------
// hash table
// ---------------------------------------------------------------------------
// One slot = 16 bytes {key, count, pad}: the key probe and the count update touch
// the same 32-byte sector, i.e. one random DRAM access (and one TLB miss) per insert
// instead of two with split key/count arrays.
struct __align__(16) HashSlot {
    unsigned long long key;
    uint32_t count;
    uint32_t pad;
};

struct HashTable {
    HashSlot* slots;
    uint64_t mask;          // capacity - 1 (capacity is a power of two)
    unsigned long long* distinct;  // number of occupied slots
    uint32_t* full;         // set when a probe sequence wrapped the whole table / load limit
    uint64_t max_distinct;
};

__device__ __forceinline__ void hash_add(const HashTable& t, uint64_t code, uint32_t add) {
    uint64_t h = kc_mix64_hd(code) & t.mask;
    for (uint64_t probes = 0; probes <= t.mask; probes++) {
        unsigned long long cur = *((volatile unsigned long long*)&t.slots[h].key);
        if (cur == KEY_EMPTY) {
            cur = atomicCAS(&t.slots[h].key, (unsigned long long)KEY_EMPTY, (unsigned long long)code);
            if (cur == KEY_EMPTY) {
                const unsigned long long d = atomicAdd(t.distinct, 1ull);
                if (d + 1 > t.max_distinct) *t.full = 1u;
                cur = code;
            }
        }
        if (cur == code) {
            atomicAdd(&t.slots[h].count, add);
            return;
        }
        h = (h + 1) & t.mask;
    }
    *t.full = 1u;
}

template <int HALO>
__global__ void __launch_bounds__(256) sparse_hash_kernel(ScanGeom g, HashTable t) {
    const int k = g.k;
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;
    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gpw = (ngroups + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t gb = g.g_begin + min(w * gpw, ngroups);
    const uint64_t ge = g.g_begin + min((w + 1) * gpw, ngroups);
    kc_warp_scan<HALO>(g, gb, ge, [&](const LaneWindow<HALO>& lw, uint64_t) {
        const uint32_t ok = lw.ok & 0xFFFFu;
        if (ok == 0) return;
        if (*((volatile uint32_t*)t.full)) return;  // table exhausted: the host retries larger
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (ok & (1u << j)) hash_add(t, lw.code64(j, kmask), 1u);
    });
}

__global__ void hash_init_kernel(HashSlot* slots, uint64_t cap) {
    const uint4 empty = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    uint4* p = reinterpret_cast<uint4*>(slots);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = empty;
}

// Compaction: per CTA iteration one global atomic reserves room for every occupied
// slot the CTA saw (warp ballots -> shared counter -> one atomicAdd), instead of one
// same-address global atomic per warp.
__global__ void __launch_bounds__(256) hash_compact_kernel(const HashSlot* __restrict__ slots, uint64_t cap,
                                                           uint64_t* __restrict__ out_keys,
                                                           uint32_t* __restrict__ out_counts,
                                                           unsigned long long* cursor) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t capr = (cap + blockDim.x - 1) / blockDim.x * blockDim.x;
    const uint4* p = reinterpret_cast<const uint4*>(slots);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capr; i += stride) {
        uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        if (i < cap) v = kc_ldg_stream(p + i);
        const bool has = !(v.x == 0xFFFFFFFFu && v.y == 0xFFFFFFFFu);
        const uint32_t m = __ballot_sync(0xffffffffu, has);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t c = s_warp[w];
                s_warp[w] = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (has) {
            const uint64_t o = s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
            out_keys[o] = ((uint64_t)v.y << 32) | v.x;
            out_counts[o] = v.z;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// sort path: one code (or the EMPTY sentinel) per window
// ---------------------------------------------------------------------------
template <int HALO>
__global__ void __launch_bounds__(256) sparse_codes_kernel(ScanGeom g, uint64_t* __restrict__ out) {
    const int k = g.k;
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;
    const uint64_t invalid = 1ull << (2 * k);  // sorts behind every code within 2k+1 key bits
    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gpw = (ngroups + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t gb = g.g_begin + min(w * gpw, ngroups);
    const uint64_t ge = g.g_begin + min((w + 1) * gpw, ngroups);
    kc_warp_scan<HALO>(g, gb, ge, [&](const LaneWindow<HALO>& lw, uint64_t a0) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const uint64_t a = a0 + j;
            if (a >= g.wlo && a < g.whi) out[a - g.wlo] = (lw.ok & (1u << j)) ? lw.code64(j, kmask) : invalid;
        }
    });
}

// trims the EMPTY run at the end of a sorted unique list
__global__ void count_valid_runs_kernel(const uint64_t* __restrict__ uniq, const unsigned long long* nruns,
                                        unsigned long long* nvalid, uint64_t invalid) {
    const unsigned long long n = *nruns;
    *nvalid = (n > 0 && uniq[n - 1] == invalid) ? n - 1 : n;
}

__global__ void owner_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t owners,
                                  unsigned long long* hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[kc_mix64_hd(keys[i]) % owners], 1ull);
}

__global__ void owner_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                     uint64_t n, uint32_t owners, unsigned long long* cursor,
                                     uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_counts) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
        const unsigned long long o = atomicAdd(&cursor[kc_mix64_hd(key) % owners], 1ull);
        out_keys[o] = key;
        out_counts[o] = counts[i];
    }
}

// ---------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------
namespace {

struct DevBuf {  // pool memory (kc_pool_alloc), returned when the buffer goes out of scope
    void* p = nullptr;
    ~DevBuf() { kc_pool_free(p); }
    cudaError_t alloc(size_t n) { return kc_pool_alloc(&p, n); }
    void reset() {
        kc_pool_free(p);
        p = nullptr;
    }
    template <typename T>
    T* as() {
        return (T*)p;
    }
    void* release() {
        void* q = p;
        p = nullptr;
        return q;
    }
};

int grid_for(kc_ctx* ctx, uint64_t ngroups) {
    uint64_t want = (ngroups + 63) / 64;
    uint64_t maxg = (uint64_t)ctx->sm_count * 8;
    return (int)(want < 1 ? 1 : (want > maxg ? maxg : want));
}

uint64_t next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// sort (keys,counts) by key and add up duplicates -> new kc_sparse
int sort_reduce_pairs(kc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n, int key_bits,
                      kc_sparse** out) {
    cudaStream_t st = ctx->stream;
    *out = nullptr;  // published only on success: an error return never leaves a live handle behind
    struct Owned {
        kc_sparse* p;
        ~Owned() { kc_sparse_free(p); }
    } own{new kc_sparse()};
    kc_sparse* res = own.p;
    res->ctx = ctx;
    res->device = ctx->device;
    if (n == 0) {
        *out = res;
        own.p = nullptr;
        return KC_OK;
    }
    DevBuf ks, cs, uk, uc, nr, tmp;
    if (ks.alloc(n * 8) || cs.alloc(n * 4) || uk.alloc(n * 8) || uc.alloc(n * 4) || nr.alloc(16)) {
        cudaGetLastError();
        return kc_set_error(ctx, KC_ERR_NOMEM, "sort_reduce_pairs: out of device memory for %llu items", (unsigned long long)n);
    }
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, d_keys, ks.as<uint64_t>(), d_counts, cs.as<uint32_t>(), (int64_t)n, 0, key_bits, st);
    cub::DeviceReduce::ReduceByKey(nullptr, t2, ks.as<uint64_t>(), uk.as<uint64_t>(), cs.as<uint32_t>(), uc.as<uint32_t>(),
                                   nr.as<unsigned long long>(), cub::Sum(), (int64_t)n, st);
    if (tmp.alloc(t1 > t2 ? t1 : t2)) {
        cudaGetLastError();
        return kc_set_error(ctx, KC_ERR_NOMEM, "sort_reduce_pairs: out of device memory for sort scratch");
    }
    size_t tb = t1 > t2 ? t1 : t2;
    KC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tb, d_keys, ks.as<uint64_t>(), d_counts, cs.as<uint32_t>(), (int64_t)n, 0,
                                                 key_bits, st));
    ctx->launches += 4;
    tb = t1 > t2 ? t1 : t2;
    KC_CUDA(ctx, cub::DeviceReduce::ReduceByKey(tmp.p, tb, ks.as<uint64_t>(), uk.as<uint64_t>(), cs.as<uint32_t>(),
                                                uc.as<uint32_t>(), nr.as<unsigned long long>(), cub::Sum(), (int64_t)n, st));
    ctx->launches += 2;
    unsigned long long nruns = 0;
    KC_CUDA(ctx, cudaMemcpyAsync(&nruns, nr.p, 8, cudaMemcpyDeviceToHost, st));
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    res->size = nruns;
    res->d_keys = (uint64_t*)uk.release();
    res->d_counts = (uint32_t*)uc.release();
    *out = res;
    own.p = nullptr;
    return KC_OK;
}

template <int HALO>
int run_hash(kc_ctx* ctx, const ScanGeom& g, uint64_t capacity, DevBuf& okeys, DevBuf& ocounts, uint64_t* ndistinct,
             bool* full) {
    cudaStream_t st = ctx->stream;
    DevBuf slots, ctl;
    if (slots.alloc(capacity * sizeof(HashSlot)) || ctl.alloc(64)) {
        cudaGetLastError();
        return kc_set_error(ctx, KC_ERR_NOMEM, "hash table of %llu slots does not fit in device memory",
                            (unsigned long long)capacity);
    }
    KC_CUDA(ctx, cudaMemsetAsync(ctl.p, 0, 64, st));
    KC_LAUNCH(hash_init_kernel, ctx->sm_count * 8, 256, 0, st, slots.as<HashSlot>(), capacity);
    KC_LAUNCH_CHECK(ctx, "hash_init_kernel");
    HashTable t;
    t.slots = slots.as<HashSlot>();
    t.mask = capacity - 1;
    t.distinct = ctl.as<unsigned long long>();
    t.full = (uint32_t*)(ctl.as<unsigned long long>() + 1);
    t.max_distinct = capacity - capacity / 4;  // load factor <= 0.75
    const uint64_t ngroups = g.g_end - g.g_begin;
    if (ngroups) {
        KC_LAUNCH(sparse_hash_kernel<HALO>, grid_for(ctx, ngroups), 256, 0, st, g, t);
        KC_LAUNCH_CHECK(ctx, "sparse_hash_kernel");
    }
    unsigned long long h[2] = {0, 0};
    KC_CUDA(ctx, cudaMemcpyAsync(h, ctl.p, 16, cudaMemcpyDeviceToHost, st));
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    *ndistinct = h[0];
    *full = (uint32_t)h[1] != 0;
    if (*full) return KC_OK;
    if (okeys.alloc(h[0] * 8) || ocounts.alloc(h[0] * 4)) {
        cudaGetLastError();
        return kc_set_error(ctx, KC_ERR_NOMEM, "out of device memory compacting %llu k-mers", h[0]);
    }
    unsigned long long* cursor = ctl.as<unsigned long long>() + 2;
    KC_LAUNCH(hash_compact_kernel, ctx->sm_count * 8, 256, 0, st, t.slots, capacity, okeys.as<uint64_t>(),
                                                           ocounts.as<uint32_t>(), cursor);
    KC_LAUNCH_CHECK(ctx, "hash_compact_kernel");
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    return KC_OK;
}

}  // namespace

extern "C" {

int kc_count_sparse(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, int algo, uint64_t capacity_hint,
                    kc_sparse** out) {
    if (!ctx || !out) return KC_ERR_INVALID;
    *out = nullptr;
    if (k < 1 || k > KC_MAX_K) return kc_set_error(ctx, KC_ERR_INVALID, "sparse k must be 1..%d, got %d", KC_MAX_K, k);
    const bool unsorted = (algo & KC_SPARSE_UNSORTED) != 0;
    const bool no_fallback = (algo & KC_SPARSE_NO_FALLBACK) != 0;
    algo &= ~(KC_SPARSE_UNSORTED | KC_SPARSE_NO_FALLBACK);
    if (algo != KC_SPARSE_HASH && algo != KC_SPARSE_SORT && algo != KC_SPARSE_RADIX && algo != KC_SPARSE_AUTO)
        return kc_set_error(ctx, KC_ERR_INVALID, "unknown sparse algo %d", algo);
    // KC_SPARSE_AUTO, by measurement on B200 (profiles/r02_sparse_*.json): the radix path beat the hash
    // table at every size and coverage tried (config 4: 135 vs 373 ms at 1/5 scale, 0.68 vs 1.79 s at full scale;
    // config 5: 221 vs 546 ms at 1/10, 319 vs 1471 ms at 1/4), so it is the default wherever it exists (2k > 20);
    // skewed inputs overflow a region and are recounted by the hash table, tiny ones go there directly (below).
    if (algo == KC_SPARSE_AUTO) algo = (2 * k > 20) ? KC_SPARSE_RADIX : KC_SPARSE_HASH;
    if (!d_data && nbytes) return kc_set_error(ctx, KC_ERR_INVALID, "null data");
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    const uint64_t nwin = nbytes >= (uint64_t)k ? nbytes - k + 1 : 0;
    if (nwin == 0) {
        kc_sparse* r = new kc_sparse();
        r->ctx = ctx;
        r->device = ctx->device;
        *out = r;
        return KC_OK;
    }
    const int halo = (k <= 17) ? 1 : 2;

    // 2^20 leaves and a 1024-partition queue only pay off on real volumes: below 4 M windows the
    // hash path is the faster one (KC_SPARSE_NO_FALLBACK insists on the radix kernels: tests)
    if (algo == KC_SPARSE_RADIX && !no_fallback && nwin < (1ull << 22) && !getenv("KC_SPARSE_RADIX_SHAPE")) algo = KC_SPARSE_HASH;
    if (algo == KC_SPARSE_RADIX) {
        int failed = 0;
        const int rc = kc_sparse_radix(ctx, d_data, nbytes, k, out, &failed);
        if (rc || !failed) return rc;
        if (no_fallback) return KC_ERR_TABLE_FULL;  // kc_last_error names the stage and what overflowed
        algo = KC_SPARSE_HASH;  // a region / leaf / run list overflowed (skewed input): recount exactly
    }

    if (algo == KC_SPARSE_HASH) {
        const ScanGeom g = kc_make_geom(d_data, nbytes, 0, nwin, k);
        uint64_t bound = nwin;
        if (k < 32 && (1ull << (2 * k)) < bound) bound = 1ull << (2 * k);
        uint64_t want = capacity_hint ? capacity_hint : bound;
        if (want > bound) want = bound;
        size_t free_b = 0, total_b = 0;
        KC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        free_b += kc_pool_idle_bytes(ctx->device);
        uint64_t cap = next_pow2(want + want / 2 + 1024);
        while (cap * 16 > free_b * 6 / 10 && cap > 1024) cap >>= 1;  // keep room for the compacted copy
        bool released = false;
        for (;;) {
            DevBuf ok, oc;
            uint64_t nd = 0;
            bool full = false;
            int rc = (halo == 1) ? run_hash<1>(ctx, g, cap, ok, oc, &nd, &full) : run_hash<2>(ctx, g, cap, ok, oc, &nd, &full);
            if (rc) return rc;
            if (!full) {
                if (unsorted) {  // hand the compacted table over as it is
                    kc_sparse* res = new kc_sparse();
                    res->ctx = ctx;
                    res->device = ctx->device;
                    res->size = nd;
                    res->d_keys = (uint64_t*)ok.release();
                    res->d_counts = (uint32_t*)oc.release();
                    *out = res;
                    return KC_OK;
                }
                rc = sort_reduce_pairs(ctx, ok.as<uint64_t>(), oc.as<uint32_t>(), nd, 2 * k, out);
                return rc;
            }
            if (cap * 2 * 16 > free_b * 8 / 10 && !released) {
                // the context's scratch (e.g. the radix path's slabs, tens of GB) is of no use here: hand it back and look again
                released = true;
                ok.reset();
                oc.reset();
                kc_scratch_release(ctx);
                KC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
                free_b += kc_pool_idle_bytes(ctx->device);
            }
            if (cap * 2 * 16 > free_b * 8 / 10)
                return kc_set_error(ctx, KC_ERR_TABLE_FULL, "hash table with %llu slots overflowed and a larger one does not fit",
                                    (unsigned long long)cap);
            cap *= 2;
        }
    }

    // sort path, chunked so that codes + sort buffers stay bounded
    const uint64_t chunk = 1ull << 28;  // windows per chunk (2 GiB of codes)
    DevBuf acc_k, acc_c;
    uint64_t acc_n = 0, acc_cap = 0, merged_n = 0;
    for (uint64_t wb = 0; wb < nwin; wb += chunk) {
        const uint64_t we = (wb + chunk < nwin) ? wb + chunk : nwin;
        const uint64_t m = we - wb;
        const ScanGeom g = kc_make_geom(d_data, nbytes, wb, we, k);
        DevBuf codes, sorted, uniq, cnts, ctl, tmp;
        if (codes.alloc(m * 8) || sorted.alloc(m * 8) || uniq.alloc(m * 8) || cnts.alloc(m * 4) || ctl.alloc(32)) {
            cudaGetLastError();
            return kc_set_error(ctx, KC_ERR_NOMEM, "sort path: out of device memory for a %llu-window chunk", (unsigned long long)m);
        }
        const uint64_t ngroups = g.g_end - g.g_begin;
        if (halo == 1)
            KC_LAUNCH(sparse_codes_kernel<1>, grid_for(ctx, ngroups), 256, 0, st, g, codes.as<uint64_t>());
        else
            KC_LAUNCH(sparse_codes_kernel<2>, grid_for(ctx, ngroups), 256, 0, st, g, codes.as<uint64_t>());
        KC_LAUNCH_CHECK(ctx, "sparse_codes_kernel");
        size_t t1 = 0, t2 = 0;
        const int sort_bits = 2 * k + 1;  // codes + the invalid marker at bit 2k
        cub::DeviceRadixSort::SortKeys(nullptr, t1, codes.as<uint64_t>(), sorted.as<uint64_t>(), (int)m, 0, sort_bits, st);
        cub::DeviceRunLengthEncode::Encode(nullptr, t2, sorted.as<uint64_t>(), uniq.as<uint64_t>(), cnts.as<uint32_t>(),
                                           ctl.as<unsigned long long>(), (int)m, st);
        size_t tb = t1 > t2 ? t1 : t2;
        if (tmp.alloc(tb)) {
            cudaGetLastError();
            return kc_set_error(ctx, KC_ERR_NOMEM, "sort path: out of device memory for sort scratch");
        }
        KC_CUDA(ctx, cub::DeviceRadixSort::SortKeys(tmp.p, tb, codes.as<uint64_t>(), sorted.as<uint64_t>(), (int)m, 0, sort_bits, st));
        tb = t1 > t2 ? t1 : t2;
        KC_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(tmp.p, tb, sorted.as<uint64_t>(), uniq.as<uint64_t>(), cnts.as<uint32_t>(),
                                                        ctl.as<unsigned long long>(), (int)m, st));
        ctx->launches += 8;
        KC_LAUNCH(count_valid_runs_kernel, 1, 1, 0, st, uniq.as<uint64_t>(), ctl.as<unsigned long long>(), ctl.as<unsigned long long>() + 1, 1ull << (2 * k));
        KC_LAUNCH_CHECK(ctx, "count_valid_runs_kernel");
        unsigned long long h[2];
        KC_CUDA(ctx, cudaMemcpyAsync(h, ctl.p, 16, cudaMemcpyDeviceToHost, st));
        KC_CUDA(ctx, cudaStreamSynchronize(st));
        const uint64_t nv = h[1];
        if (acc_n + nv > acc_cap) {
            uint64_t ncap = (acc_n + nv) * 2;
            DevBuf nk, nc;
            if (nk.alloc(ncap * 8) || nc.alloc(ncap * 4)) {
                cudaGetLastError();
                return kc_set_error(ctx, KC_ERR_NOMEM, "sort path: out of device memory accumulating runs");
            }
            if (acc_n) {
                KC_CUDA(ctx, cudaMemcpyAsync(nk.p, acc_k.p, acc_n * 8, cudaMemcpyDeviceToDevice, st));
                KC_CUDA(ctx, cudaMemcpyAsync(nc.p, acc_c.p, acc_n * 4, cudaMemcpyDeviceToDevice, st));
                KC_CUDA(ctx, cudaStreamSynchronize(st));
            }
            std::swap(acc_k.p, nk.p);
            std::swap(acc_c.p, nc.p);
            acc_cap = ncap;
        }
        KC_CUDA(ctx, cudaMemcpyAsync(acc_k.as<uint64_t>() + acc_n, uniq.p, nv * 8, cudaMemcpyDeviceToDevice, st));
        KC_CUDA(ctx, cudaMemcpyAsync(acc_c.as<uint32_t>() + acc_n, cnts.p, nv * 4, cudaMemcpyDeviceToDevice, st));
        KC_CUDA(ctx, cudaStreamSynchronize(st));
        acc_n += nv;
        // Keep the accumulated runs bounded: with deep coverage most runs of different
        // chunks are the same k-mers, so merge (sort + reduce-by-key) whenever the list
        // has grown by more than a chunk's worth since the last merge.
        if (acc_n > merged_n + 3 * chunk && wb + chunk < nwin) {
            codes.reset();  // the chunk's buffers are no longer needed: make room for the merge
            sorted.reset();
            uniq.reset();
            cnts.reset();
            tmp.reset();
            kc_sparse* part = nullptr;
            int rc = sort_reduce_pairs(ctx, acc_k.as<uint64_t>(), acc_c.as<uint32_t>(), acc_n, 2 * k, &part);
            if (rc) {
                kc_sparse_free(part);
                return rc;
            }
            kc_pool_free(acc_k.p);
            kc_pool_free(acc_c.p);
            acc_k.p = part->d_keys;
            acc_c.p = part->d_counts;
            acc_n = merged_n = part->size;
            acc_cap = part->size;  // the arrays hold `n input items` each, at least size
            part->d_keys = nullptr;
            part->d_counts = nullptr;
            kc_sparse_free(part);
        }
    }
    return sort_reduce_pairs(ctx, acc_k.as<uint64_t>(), acc_c.as<uint32_t>(), acc_n, 2 * k, out);
}

void kc_sparse_free(kc_sparse* s) {
    if (!s) return;
    if (s->device >= 0) {  // not s->ctx->device: the ctx may have been destroyed before its results
        DeviceGuard dg(s->device);
        kc_pool_free(s->d_keys);
        kc_pool_free(s->d_counts);
    }
    delete s;
}
uint64_t kc_sparse_size(const kc_sparse* s) { return s ? s->size : 0; }
const uint64_t* kc_sparse_d_keys(const kc_sparse* s) { return s ? s->d_keys : nullptr; }
const uint32_t* kc_sparse_d_counts(const kc_sparse* s) { return s ? s->d_counts : nullptr; }

int kc_sparse_copy_to_host(kc_ctx* ctx, const kc_sparse* s, uint64_t* h_keys, uint32_t* h_counts) {
    if (!ctx || !s) return KC_ERR_INVALID;
    if (s->size == 0) return KC_OK;
    DeviceGuard dg(ctx->device);
    if (h_keys) KC_CUDA(ctx, cudaMemcpyAsync(h_keys, s->d_keys, s->size * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_counts) KC_CUDA(ctx, cudaMemcpyAsync(h_counts, s->d_counts, s->size * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

int kc_sparse_bucket_by_owner(kc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n,
                              uint32_t num_owners, uint64_t* d_keys_out, uint32_t* d_counts_out,
                              uint64_t* h_bucket_sizes) {
    if (!ctx || !h_bucket_sizes || num_owners == 0 || num_owners > 1024) return kc_set_error(ctx, KC_ERR_INVALID, "bucket_by_owner: bad argument");
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    for (uint32_t i = 0; i < num_owners; i++) h_bucket_sizes[i] = 0;
    if (n == 0) return KC_OK;
    DevBuf ctl;
    if (ctl.alloc(2 * 8 * (size_t)num_owners)) {
        cudaGetLastError();
        return kc_set_error(ctx, KC_ERR_NOMEM, "bucket_by_owner: out of device memory");
    }
    unsigned long long* hist = ctl.as<unsigned long long>();
    unsigned long long* cursor = hist + num_owners;
    KC_CUDA(ctx, cudaMemsetAsync(hist, 0, 8 * (size_t)num_owners, st));
    const int grid = ctx->sm_count * 8;
    KC_LAUNCH(owner_hist_kernel, grid, 256, 0, st, d_keys, n, num_owners, hist);
    KC_LAUNCH_CHECK(ctx, "owner_hist_kernel");
    std::vector<unsigned long long> h(num_owners), c(num_owners);
    KC_CUDA(ctx, cudaMemcpyAsync(h.data(), hist, 8 * (size_t)num_owners, cudaMemcpyDeviceToHost, st));
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    unsigned long long run = 0;
    for (uint32_t i = 0; i < num_owners; i++) {
        c[i] = run;
        run += h[i];
        h_bucket_sizes[i] = h[i];
    }
    KC_CUDA(ctx, cudaMemcpyAsync(cursor, c.data(), 8 * (size_t)num_owners, cudaMemcpyHostToDevice, st));
    KC_LAUNCH(owner_scatter_kernel, grid, 256, 0, st, d_keys, d_counts, n, num_owners, cursor, d_keys_out, d_counts_out);
    KC_LAUNCH_CHECK(ctx, "owner_scatter_kernel");
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    return KC_OK;
}

int kc_sparse_merge(kc_ctx* ctx, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t n, kc_sparse** out) {
    if (!ctx || !out) return KC_ERR_INVALID;
    *out = nullptr;
    if (n && (!d_keys || !d_counts)) return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_merge: null pointer");
    DeviceGuard dg(ctx->device);
    return sort_reduce_pairs(ctx, d_keys, d_counts, n, 64, out);
}

}  // extern "C"

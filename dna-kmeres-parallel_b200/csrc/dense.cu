// dense.cu — dense 4^k histogram counting (k <= 16), the hot path.
//
// Replaces the reference's count kernel sumKmereCoincidencesGlobalMemory
// (kernels.h:113-144: one CTA per sequence, one thread per 3-mer, each thread
// re-scanning the whole sequence byte by byte) with single-pass kernels:
//
//   dense_direct_kernel<SMEM>  k <= 7 : CTA-private 4^k uint32 bins in shared
//                                       memory, flushed once with global REDs.
//   dense_direct_kernel<GLOBAL> any k : one global RED per window (bins in L2).
//   partition path (k = 12)           : pass 1 groups A=5 consecutive windows
//       into one 32-bit "record" (the 16 bases they span), radix-partitions
//       records by 11 bits that all five windows share, staging them in shared
//       memory so every global write is a coalesced run; pass 2 gives each CTA
//       one partition and counts its five 8192-bin sub-tables with shared-memory
//       atomics, then adds them to the global table in 128-byte runs.
//       Random scatter to a 64 MiB table moves from L2 (<= 1 sector/clk/SM) into
//       shared memory (32 banks/clk/SM).  See DESIGN.md §3.3.
#include "common.cuh"

// ---------------------------------------------------------------------------
// direct kernels
// ---------------------------------------------------------------------------
enum { BINS_SMEM = 0, BINS_GLOBAL = 1 };

template <int MODE>
__global__ void __launch_bounds__(256) dense_direct_kernel(ScanGeom g, uint32_t* __restrict__ table) {
    extern __shared__ uint32_t s_bins[];
    const int k = g.k;
    const uint32_t nbins = (k >= 16) ? 0u : (1u << (2 * k));  // only the smem mode (k <= 7) uses it
    const uint32_t kmask = (k >= 16) ? 0xFFFFFFFFu : (nbins - 1u);
    if (MODE == BINS_SMEM) {
        for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0;
        __syncthreads();
    }
    uint32_t* bins = (MODE == BINS_SMEM) ? s_bins : table;

    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gpw = (ngroups + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t gb = g.g_begin + min(w * gpw, ngroups);
    const uint64_t ge = g.g_begin + min((w + 1) * gpw, ngroups);

    kc_warp_scan<1>(g, gb, ge, [&](const LaneWindow<1>& lw, uint64_t) {
        const uint32_t ok = lw.ok & 0xFFFFu;
        if (ok == 0) return;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (ok & (1u << j)) atomicAdd(&bins[lw.code32(j, kmask)], 1u);
        }
    });

    if (MODE == BINS_SMEM) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) {
            const uint32_t v = s_bins[i];
            if (v) atomicAdd(&table[i], v);
        }
    }
}

// ---------------------------------------------------------------------------
// partition path
// ---------------------------------------------------------------------------
template <int K_, int A_, int KB_, int CAP_>
struct PartCfg {
    static constexpr int K = K_;            // window length
    static constexpr int A = A_;            // windows per record
    static constexpr int KB = KB_;          // partition key bits
    static constexpr int CAP = CAP_;        // staged records per partition per flush
    static constexpr int P = 1 << KB;       // partitions
    static constexpr int REC_BASES = K + A - 1;
    static constexpr int KB0 = 2 * K - KB;  // key = record bits [KB0, 2K)
    static constexpr int SUB = 1 << KB0;    // bins per alignment sub-table
    static constexpr int NBINS = A * SUB;   // bins per partition (pass 2 smem)
    static constexpr uint32_t AMASK = (1u << A) - 1u;
    static_assert(REC_BASES <= 16, "record must fit 32 bits");
    static_assert(KB0 >= 2 * (A - 1), "key must lie in the bases all A windows share");
};

// bin (global table index) of the r-th window of a record
template <typename C>
__device__ __forceinline__ uint32_t part_code(uint32_t rec, int r) {
    return (rec >> (2 * r)) & ((C::K == 16) ? 0xFFFFFFFFu : ((1u << (2 * C::K)) - 1u));
}

template <typename C>
__device__ __forceinline__ void part_fallback(uint32_t rec, uint32_t okbits, uint32_t* table) {
#pragma unroll
    for (int r = 0; r < C::A; r++)
        if (okbits & (1u << r)) atomicAdd(&table[part_code<C>(rec, r)], 1u);
}

// Pass 1.  1024 threads, one CTA per SM, each warp owns a contiguous run of
// 512-byte groups.  Shared memory: cnt[P], gbase[P], buf[P][CAP].
template <typename C, int FLUSH_EVERY>
__global__ void __launch_bounds__(1024, 1)
part_scatter_kernel(ScanGeom g, uint32_t* __restrict__ table, uint32_t* __restrict__ slabs,
                    uint32_t* __restrict__ gcursor, uint32_t slab_cap) {
    extern __shared__ uint32_t smem[];
    uint32_t* cnt = smem;
    uint32_t* gbase = smem + C::P;
    uint32_t* buf = smem + 2 * C::P;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int b = tid; b < C::P; b += 1024) cnt[b] = 0;
    __syncthreads();

    // every warp of the CTA runs the same number of steps so the CTA-wide
    // flush barriers line up; warps past the end scan empty ranges.
    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * 32;
    const uint64_t gpw = (ngroups + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * 32 + (tid >> 5);
    const uint64_t gb = g.g_begin + min(w * gpw, ngroups);
    const uint64_t ge = g.g_begin + min((w + 1) * gpw, ngroups);

    auto flush = [&]() {
        __syncthreads();
        for (int b = tid; b < C::P; b += 1024) {
            const uint32_t c = min(cnt[b], (uint32_t)C::CAP);
            gbase[b] = c ? atomicAdd(&gcursor[b], c) : 0u;
        }
        __syncthreads();
        const int hw = tid >> 4, l16 = tid & 15;
        for (int b = hw; b < C::P; b += 64) {
            const uint32_t c = min(cnt[b], (uint32_t)C::CAP);
            const uint32_t gbs = gbase[b];
            for (uint32_t s = l16; s < c; s += 16) {
                const uint32_t rec = buf[b * C::CAP + s];
                const uint32_t gi = gbs + s;
                if (gi < slab_cap)
                    slabs[(uint64_t)b * slab_cap + gi] = rec;
                else
                    part_fallback<C>(rec, C::AMASK, table);
            }
        }
        __syncthreads();
        for (int b = tid; b < C::P; b += 1024) cnt[b] = 0;
        __syncthreads();
    };

    Decoded16 cur = kc_load_block(g, gb * 32 + lane);
    for (uint64_t step = 0; step < gpw; step++) {
        const uint64_t grp = gb + step;
        if (grp < ge) {  // warp-uniform
            Decoded16 nxt;
            if (grp + 1 < ge || lane < 1) {
                nxt = kc_load_block(g, (grp + 1) * 32 + lane);
            } else {
                nxt.packed = 0;
                nxt.bad = 0xFFFFu;
            }
            uint32_t p1 = __shfl_down_sync(0xffffffffu, cur.packed, 1);
            uint32_t b1 = __shfl_down_sync(0xffffffffu, cur.bad, 1);
            const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
            const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
            if (lane == 31) {
                p1 = n0p;
                b1 = n0b;
            }
            const uint32_t p0 = cur.packed;
            const uint64_t B = (uint64_t)cur.bad | ((uint64_t)b1 << 16) | (0xFFFFull << 32);
            const uint64_t a0 = (grp * 32 + lane) << 4;
            uint32_t ok = ~(uint32_t)kc_window_bad(B, C::K);  // 32 window starts
            if ((grp << 9) < g.wlo || ((grp + 1) << 9) + 16 > g.whi) {
                const int l = a0 < g.wlo ? (int)min((uint64_t)32, g.wlo - a0) : 0;
                const int h = a0 + 32 > g.whi ? (int)(g.whi > a0 ? g.whi - a0 : 0) : 32;
                const uint32_t hm = h >= 32 ? 0xFFFFFFFFu : ((1u << h) - 1u);
                const uint32_t lm = l >= 32 ? 0xFFFFFFFFu : ((1u << l) - 1u);
                ok &= hm & ~lm;
            }
            // records start at aligned coordinates wlo + A*m
            int j0;
            if (a0 >= g.wlo) {
                const uint32_t d = (uint32_t)((a0 - g.wlo) % (uint64_t)C::A);
                j0 = d ? C::A - (int)d : 0;
            } else {
                j0 = (int)min((uint64_t)64, g.wlo - a0);
            }
#pragma unroll
            for (int t = 0; t < (16 + C::A - 1) / C::A; t++) {
                const int j = j0 + t * C::A;
                if (j < 16) {
                    const uint32_t okr = (ok >> j) & C::AMASK;
                    if (okr) {
                        const uint32_t rec = __funnelshift_r(p0, p1, 2 * j);
                        if (okr == C::AMASK) {
                            const uint32_t pid = (rec >> C::KB0) & (C::P - 1);
                            const uint32_t slot = atomicAdd(&cnt[pid], 1u);
                            if (slot < (uint32_t)C::CAP) {
                                buf[pid * C::CAP + slot] = rec;
                            } else {
                                const uint32_t gi = atomicAdd(&gcursor[pid], 1u);
                                if (gi < slab_cap)
                                    slabs[(uint64_t)pid * slab_cap + gi] = rec;
                                else
                                    part_fallback<C>(rec, C::AMASK, table);
                            }
                        } else {
                            part_fallback<C>(rec, okr, table);
                        }
                    }
                }
            }
            cur = nxt;
        }
        if ((step % FLUSH_EVERY) == FLUSH_EVERY - 1) flush();
    }
    flush();
}

// Pass 2.  One partition at a time per CTA (dynamic queue); NBINS uint32 bins
// in shared memory.
template <typename C>
__global__ void __launch_bounds__(1024, 1)
part_count_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slabs,
                  const uint32_t* __restrict__ gcursor, uint32_t slab_cap,
                  uint32_t* __restrict__ work_counter) {
    extern __shared__ uint32_t bins[];
    __shared__ uint32_t s_part;
    const int tid = threadIdx.x;
    constexpr uint32_t SUBMASK = C::SUB - 1;
    for (;;) {
        if (tid == 0) s_part = atomicAdd(work_counter, 1u);
        {
            uint4* b4 = reinterpret_cast<uint4*>(bins);
            for (int i = tid; i < C::NBINS / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= (uint32_t)C::P) break;
        const uint32_t n = min(gcursor[part], slab_cap);
        const uint32_t* src = slabs + (uint64_t)part * slab_cap;  // slab_cap % 4 == 0 -> 16 B aligned
        auto count_rec = [&](uint32_t rec) {
            // Y = record with the key bits removed; window r's bin = Y bits [2r, 2r+KB0)
            const uint32_t Y = (rec & SUBMASK) | ((rec >> (2 * C::K)) << C::KB0);
#pragma unroll
            for (int r = 0; r < C::A; r++) atomicAdd(&bins[r * C::SUB + ((Y >> (2 * r)) & SUBMASK)], 1u);
        };
        const uint32_t n4 = n >> 2;
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
        for (uint32_t i = tid; i < n4; i += 1024) {
            const uint4 v = kc_ldg_stream(src4 + i);
            count_rec(v.x);
            count_rec(v.y);
            count_rec(v.z);
            count_rec(v.w);
        }
        for (uint32_t i = (n4 << 2) + tid; i < n; i += 1024) count_rec(src[i]);
        __syncthreads();
        // add the sub-tables to the global table: for alignment r the bin is
        //   low (KB0-2r bits) | key << (KB0-2r) | high (2r bits) << (2K-2r)
        for (int idx = tid; idx < C::NBINS; idx += 1024) {
            const uint32_t v = bins[idx];
            if (v) {
                const int r = idx / C::SUB;
                const uint32_t f = idx & SUBMASK;
                const int lowbits = C::KB0 - 2 * r;
                const uint32_t low = f & ((1u << lowbits) - 1u);
                const uint32_t high = f >> lowbits;
                const uint32_t code = low | (part << lowbits) | (high << (2 * C::K - 2 * r));
                atomicAdd(&table[code], v);
            }
        }
        __syncthreads();
    }
}

using Part12 = PartCfg<12, 5, 11, 24>;
constexpr int kPart12Flush = 6;

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int dense_direct(kc_ctx* ctx, const ScanGeom& g, uint32_t* d_table, cudaStream_t st) {
    const uint64_t ngroups = g.g_end - g.g_begin;
    if (ngroups == 0) return KC_OK;
    const int k = g.k;
    const bool use_smem = (k <= 7);
    // 8 warps per CTA; aim at >= 8 groups per warp, at most 8 CTAs per SM
    uint64_t want = (ngroups + 63) / 64;
    uint64_t maxg = (uint64_t)ctx->sm_count * (use_smem && k == 7 ? 3 : 8);
    int grid = (int)(want < 1 ? 1 : (want > maxg ? maxg : want));
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
    if (use_smem) {
        const size_t smem = sizeof(uint32_t) << (2 * k);
        if (smem > 48 * 1024)
            KC_CUDA(ctx, cudaFuncSetAttribute(dense_direct_kernel<BINS_SMEM>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_direct_kernel<BINS_SMEM><<<grid, 256, smem, st>>>(g, d_table);
    } else {
        dense_direct_kernel<BINS_GLOBAL><<<grid, 256, 0, st>>>(g, d_table);
    }
    KC_LAUNCH_CHECK(ctx, "dense_direct_kernel");
    if (ctx->timing) {
        KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
        ctx->timed_kernels = 1;
    }
    return KC_OK;
}

template <typename C, int FLUSH>
static int dense_partition(kc_ctx* ctx, const ScanGeom& g, uint32_t* d_table, cudaStream_t st) {
    const uint64_t ngroups = g.g_end - g.g_begin;
    if (ngroups == 0) return KC_OK;
    const uint64_t nwin = g.whi - g.wlo;
    const uint64_t nrec = (nwin + C::A - 1) / C::A;
    // slab capacity: mean + 12.5 % + slack, multiple of 4 records (16 B)
    uint64_t cap = nrec / C::P;
    cap = cap + cap / 8 + 4096;
    cap = (cap + 3) & ~3ull;
    if (cap > 0xFFFFFFF0ull) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition slab too large");
    const size_t slab_bytes = (size_t)cap * C::P * sizeof(uint32_t);
    const size_t ctl_bytes = (C::P + 64) * sizeof(uint32_t);
    int rc = kc_scratch_reserve(ctx, slab_bytes + ctl_bytes);
    if (rc) return rc;
    uint32_t* gcursor = (uint32_t*)ctx->scratch;
    uint32_t* work = gcursor + C::P;
    uint32_t* slabs = gcursor + C::P + 64;
    KC_CUDA(ctx, cudaMemsetAsync(gcursor, 0, ctl_bytes, st));

    const size_t smem1 = (size_t)(2 * C::P + C::P * C::CAP) * sizeof(uint32_t);
    const size_t smem2 = (size_t)C::NBINS * sizeof(uint32_t);
    KC_CUDA(ctx, cudaFuncSetAttribute(part_scatter_kernel<C, FLUSH>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    KC_CUDA(ctx, cudaFuncSetAttribute(part_count_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem2));
    uint64_t want = (ngroups + 31) / 32;
    int grid1 = (int)(want > (uint64_t)ctx->sm_count ? (uint64_t)ctx->sm_count : want);
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
    part_scatter_kernel<C, FLUSH><<<grid1, 1024, smem1, st>>>(g, d_table, slabs, gcursor, (uint32_t)cap);
    KC_LAUNCH_CHECK(ctx, "part_scatter_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
    int grid2 = ctx->sm_count < C::P ? ctx->sm_count : C::P;
    part_count_kernel<C><<<grid2, 1024, smem2, st>>>(d_table, slabs, gcursor, (uint32_t)cap, work);
    KC_LAUNCH_CHECK(ctx, "part_count_kernel");
    if (ctx->timing) {
        KC_CUDA(ctx, cudaEventRecord(ctx->tev[2], st));
        ctx->timed_kernels = 2;
    }
    return KC_OK;
}

static uint64_t g_partition_min_windows = 1ull << 24;  // below this the direct path wins

extern "C" int kc_count_dense_range_async(kc_ctx* ctx, const char* d_data, uint64_t nbytes,
                                          uint64_t win_begin, uint64_t win_end, int k,
                                          uint32_t* d_table, int algo, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!d_table || (!d_data && nbytes)) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    if (algo != KC_DENSE_AUTO && algo != KC_DENSE_DIRECT && algo != KC_DENSE_PARTITION)
        return kc_set_error(ctx, KC_ERR_INVALID, "unknown dense algo %d", algo);
    DeviceGuard dg(ctx->device);
    if (nbytes < (uint64_t)k) return KC_OK;
    const uint64_t nwin = nbytes - k + 1;
    if (win_end > nwin) win_end = nwin;
    if (win_begin >= win_end) return KC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, k);
    const bool can_part = (k == 12);
    if (algo == KC_DENSE_PARTITION && !can_part)
        return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition path is built for k=12 only (k=%d)", k);
    const bool use_part =
        can_part && (algo == KC_DENSE_PARTITION ||
                     (algo == KC_DENSE_AUTO && (win_end - win_begin) >= g_partition_min_windows));
    if (use_part) return dense_partition<Part12, kPart12Flush>(ctx, g, d_table, st);
    return dense_direct(ctx, g, d_table, st);
}

extern "C" int kc_count_dense_async(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k,
                                    uint32_t* d_table, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!d_table) return kc_set_error(ctx, KC_ERR_INVALID, "null table");
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaMemsetAsync(d_table, 0, sizeof(uint32_t) << (2 * k), (cudaStream_t)stream));
    return kc_count_dense_range_async(ctx, d_data, nbytes, 0, nbytes, k, d_table, KC_DENSE_AUTO, stream);
}

extern "C" int kc_count_dense(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, uint32_t* d_table) {
    if (!ctx) return KC_ERR_INVALID;
    int rc = kc_count_dense_async(ctx, d_data, nbytes, k, d_table, ctx->stream);
    if (rc) return rc;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

// End to end from host memory: chunked H2D into two device buffers on the copy
// stream, counting on the compute stream, events for the hand-off.  Chunk c is
// counted for the windows that END inside it (so only bytes already on the
// device are touched); the device buffer holds the whole input.
extern "C" int kc_count_dense_host(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k,
                                   uint32_t* h_table) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!h_table || (!h_data && nbytes)) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    DeviceGuard dg(ctx->device);
    const size_t table_bytes = sizeof(uint32_t) << (2 * k);
    int rc = kc_scratch2_reserve(ctx, table_bytes + nbytes + 64);
    if (rc) return rc;
    uint32_t* d_table = (uint32_t*)ctx->scratch2;
    char* d_data = (char*)ctx->scratch2 + table_bytes;
    KC_CUDA(ctx, cudaMemsetAsync(d_table, 0, table_bytes, ctx->stream));
    const uint64_t nwin = nbytes >= (uint64_t)k ? nbytes - k + 1 : 0;
    const uint64_t chunk = 256ull << 20;
    const int nchunks = (int)((nbytes + chunk - 1) / chunk);
    cudaEvent_t ev;
    KC_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    uint64_t counted = 0;  // windows [0, counted) are done
    for (int c = 0; c < nchunks; c++) {
        const uint64_t b = (uint64_t)c * chunk;
        const uint64_t e = (b + chunk < nbytes) ? b + chunk : nbytes;
        cudaError_t ce = cudaMemcpyAsync(d_data + b, h_data + b, e - b, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev, ctx->copy_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, ev, 0);
        if (ce != cudaSuccess) {
            cudaEventDestroy(ev);
            return kc_set_error(ctx, KC_ERR_CUDA, "host staging failed: %s", cudaGetErrorString(ce));
        }
        // windows that end before byte e: start < e - k + 1
        uint64_t upto = (e >= (uint64_t)k) ? e - k + 1 : 0;
        if (upto > nwin) upto = nwin;
        if (upto > counted) {
            rc = kc_count_dense_range_async(ctx, d_data, e, counted, upto, k, d_table, KC_DENSE_AUTO, ctx->stream);
            if (rc) {
                cudaEventDestroy(ev);
                return rc;
            }
            counted = upto;
        }
    }
    cudaEventDestroy(ev);
    KC_CUDA(ctx, cudaMemcpyAsync(h_table, d_table, table_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

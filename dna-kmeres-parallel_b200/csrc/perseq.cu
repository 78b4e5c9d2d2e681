// perseq.cu — reference-shaped per-sequence count table  sums[4^k][num_seqs].
//
// Replaces sumKmereCoincidencesGlobalMemory (kernels.h:113-144, launched once
// with 54018 CTAs of 64 threads at main.cu:290, where only num_seqs CTAs work and
// a single long sequence runs on 2 warps).  Here the concatenated byte stream is
// cut into 32 KiB tiles that persistent CTAs pick up round-robin; a tile walks
// the sequences it overlaps, and each (sequence, tile) segment is scanned by all
// 8 warps of the CTA with the same WarpScanner as the dense path.  Window range
// per sequence is [off[e], off[e] + L - k + 1) with L = off[e+1]-off[e]-1, i.e.
// exactly the reference's `i < entryLength - 3` loop (kernels.h:124,133) for any
// k: the separator byte is never part of a window.
//
// Bins: k <= 6 -> CTA-private int32 table in shared memory, flushed to
// sums[bin*num_seqs + e] per segment (segments much smaller than the table
// go straight to global atomics); k >= 7 -> global atomics.
#include "common.cuh"

constexpr int PS_TILE_BYTES = 32 * 1024;
constexpr int PS_SMEM_MAX_K = 6;

__device__ __forceinline__ uint32_t ps_seq_of(const int64_t* __restrict__ off, uint32_t num_seqs, int64_t pos) {
    // largest e with off[e] <= pos   (off[0] <= pos is guaranteed by the caller)
    uint32_t lo = 0, hi = num_seqs;  // answer in [lo, hi)
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] <= pos)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
perseq_kernel(const char* __restrict__ data, const int64_t* __restrict__ off, uint32_t num_seqs, int k,
              int32_t* __restrict__ sums, int use_smem) {
    KC_DYN_SMEM(uint32_t, s_bins);
    const uint32_t nbins = (k >= 16) ? 0u : (1u << (2 * k));  // smem mode (k <= 6) only
    const uint32_t kmask = (k >= 16) ? 0xFFFFFFFFu : (nbins - 1u);
    const int warp = threadIdx.x >> 5;
    const int64_t begin = off[0];
    const int64_t total = off[num_seqs];
    const int64_t ntiles = (total - begin + PS_TILE_BYTES - 1) / PS_TILE_BYTES;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) s_bins[i] = 0;
        __syncthreads();
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t t0 = begin + tile * PS_TILE_BYTES;
        const int64_t t1 = min(t0 + (int64_t)PS_TILE_BYTES, total);
        for (uint32_t e = ps_seq_of(off, num_seqs, t0); e < num_seqs; e++) {
            const int64_t s = off[e];
            if (s >= t1) break;
            const int64_t L = off[e + 1] - s - 1;  // kernels.h:124 minus the separator
            if (L < k) continue;
            const int64_t wb = max(s, t0);
            const int64_t we = min(s + L - k + 1, t1);
            if (wb >= we) continue;
            // bytes of this sequence only: [.., s+L) readable, windows in [wb, we)
            const ScanGeom g = kc_make_geom(data, (uint64_t)(s + L), (uint64_t)wb, (uint64_t)we, k);
            const uint64_t ng = g.g_end - g.g_begin;
            const uint64_t per = (ng + 7) >> 3;
            const uint64_t gb = g.g_begin + min((uint64_t)warp * per, ng);
            const uint64_t ge = g.g_begin + min((uint64_t)(warp + 1) * per, ng);
            const bool seg_smem = use_smem && (uint64_t)(we - wb) * 2 >= nbins;  // CTA-uniform
            if (seg_smem) {
                kc_warp_scan<1>(g, gb, ge, [&](const LaneWindow<1>& lw, uint64_t) {
                    const uint32_t ok = lw.ok & 0xFFFFu;
                    if (ok == 0) return;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (ok & (1u << j)) atomicAdd(&s_bins[lw.code32(j, kmask)], 1u);
                });
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) {
                    const uint32_t v = s_bins[i];
                    if (v) {
                        atomicAdd(&sums[(uint64_t)i * num_seqs + e], (int32_t)v);
                        s_bins[i] = 0;
                    }
                }
                __syncthreads();
            } else {
                kc_warp_scan<1>(g, gb, ge, [&](const LaneWindow<1>& lw, uint64_t) {
                    const uint32_t ok = lw.ok & 0xFFFFu;
                    if (ok == 0) return;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (ok & (1u << j))
                            atomicAdd(&sums[(uint64_t)lw.code32(j, kmask) * num_seqs + e], 1);
                });
            }
        }
    }
}

extern "C" int kc_count_per_seq_async(kc_ctx* ctx, const char* d_data, const int64_t* d_offsets,
                                      uint32_t num_seqs, int k, int32_t* d_sums, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "per-seq k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!d_sums || !d_offsets) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    DeviceGuard dg(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (num_seqs == 0) return KC_OK;
    const size_t bytes = ((size_t)sizeof(int32_t) << (2 * k)) * num_seqs;
    KC_CUDA(ctx, cudaMemsetAsync(d_sums, 0, bytes, st));
    const int use_smem = k <= PS_SMEM_MAX_K;
    const size_t smem = use_smem ? (sizeof(uint32_t) << (2 * k)) : 0;
    const int grid = ctx->sm_count * 4;
    KC_LAUNCH(perseq_kernel, grid, 256, smem, st, d_data, d_offsets, num_seqs, k, d_sums, use_smem);
    KC_LAUNCH_CHECK(ctx, "perseq_kernel");
    return KC_OK;
}

extern "C" int kc_count_per_seq(kc_ctx* ctx, const char* d_data, const int64_t* d_offsets, uint32_t num_seqs,
                                int k, int32_t* d_sums) {
    if (!ctx) return KC_ERR_INVALID;
    int rc = kc_count_per_seq_async(ctx, d_data, d_offsets, num_seqs, k, d_sums, ctx->stream);
    if (rc) return rc;
    DeviceGuard dg(ctx->device);
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

// ---------------------------------------------------------------------------
// "next" row f1: all-pairs k-mer distance (minKmeres2 kernels.h:85-109,
// sequentialKmerCount2 main.cu:587-621).  One launch instead of num_seqs
// synchronous ones (main.cu:327-335).  Row i of the table is staged in shared
// memory in chunks (the reference stages all 64 bins, kernels.h:86,91-94);
// threads own the later sequences j, so reads sums[j + n*p] are coalesced.
// The sum of minima is accumulated in int64 like the CPU reference's `long`
// (main.cu:593,607-613) and converted once; the reference GPU kernel adds in
// float (kernels.h:97,104), identical while the sum stays below 2^24.
// ---------------------------------------------------------------------------
constexpr int DIST_CHUNK = 1024;

__global__ void __launch_bounds__(256)
distance_kernel(const int32_t* __restrict__ sums, const int64_t* __restrict__ off, uint32_t n, int k,
                float* __restrict__ dist) {
    __shared__ int32_t row[DIST_CHUNK];
    const uint32_t i = blockIdx.y;
    const uint32_t j = i + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 + blockIdx.x * blockDim.x >= n) return;  // whole CTA idle (uniform)
    const uint64_t nbins = 1ull << (2 * k);
    long long acc = 0;
    for (uint64_t p0 = 0; p0 < nbins; p0 += DIST_CHUNK) {
        const int m = (int)min((uint64_t)DIST_CHUNK, nbins - p0);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += blockDim.x) row[t] = sums[(p0 + t) * n + i];
        __syncthreads();
        if (j < n) {
            for (int t = 0; t < m; t++) acc += min(row[t], sums[(p0 + t) * n + j]);
        }
    }
    if (j < n) {
        const long long Li = off[i + 1] - off[i] - 1, Lj = off[j + 1] - off[j] - 1;
        const long long minLength = Li < Lj ? Li : Lj;
        const float d = 1 - (float)acc / (minLength - k + 1);
        const long long ii = i + 1, gap = j - i, nn = n;  // kernels.h:46-48 index, 1-based i
        dist[(nn * (ii - 1) - (((ii - 2) * (ii - 1)) / 2)) + (gap - ii)] = d;
    }
}

extern "C" int kc_kmer_distance(kc_ctx* ctx, const int32_t* d_sums, const int64_t* d_offsets, uint32_t num_seqs,
                                int k, float* d_dist) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "distance k must be 1..%d", KC_MAX_DENSE_K);
    if (!d_sums || !d_offsets || (!d_dist && num_seqs > 1)) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    if (num_seqs < 2) return KC_OK;
    if (num_seqs > 65535) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "distance step supports up to 65535 sequences");
    DeviceGuard dg(ctx->device);
    dim3 grid((num_seqs - 1 + 255) / 256, num_seqs - 1);
    KC_LAUNCH(distance_kernel, grid, 256, 0, ctx->stream, d_sums, d_offsets, num_seqs, k, d_dist);
    KC_LAUNCH_CHECK(ctx, "distance_kernel");
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

// check.cu — full-scale self-checks of a count, independent of the counting kernels.
//
// A k-mer count is a multiset identity: the counted (code, count) pairs must hold exactly the
// valid windows of the input.  With h = mix64 (the finaliser shared with the oracle),
//   F_in  = sum over the valid windows w of the input   h(code(w))            (mod 2^64)
//   F_out = sum over the distinct k-mers c of a result   count(c) * h(c)       (mod 2^64)
// are equal for a correct count, and a lost, duplicated or altered window changes F_out by a
// 64-bit pseudo-random amount.  F_in is ONE streaming scan (no table, no atomics on the data
// path), additive over any partition of the input (shards, ranks), so it checks a 15 Gbp
// config-4 count or an 8-GPU count at full scale where the CPU oracle cannot go; the sums of
// counts give "sum of counts == valid windows" on the way.  bench.py reports both in its
// result line; tests/ compare kc_window_fingerprint itself with the oracle.
//
// Window semantics are the reference's (main.cu:636-646): k consecutive bytes, counted iff all
// are upper-case ACGT; code = sum code(s[p]) * 4^p (utils.h:30-47).
#include "common.cuh"

namespace {

__device__ __forceinline__ void block_add_u64(unsigned long long v, unsigned long long c, unsigned long long* out) {
    // warp shuffle reduce, then one atomic pair per warp
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        v += __shfl_down_sync(0xffffffffu, (unsigned long long)v, d);
        c += __shfl_down_sync(0xffffffffu, (unsigned long long)c, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, v);
        atomicAdd(out + 1, c);
    }
}

template <int HALO>
__global__ void __launch_bounds__(256) window_fp_kernel(ScanGeom g, unsigned long long* __restrict__ out) {
    const int k = g.k;
    const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t gb = g.g_begin + w * ngroups / nwarps, ge = g.g_begin + (w + 1) * ngroups / nwarps;
    unsigned long long acc = 0, cnt = 0;
    kc_warp_scan<HALO>(g, gb, ge, [&](const LaneWindow<HALO>& lw, uint64_t) {
        uint32_t m = lw.ok & 0xFFFFu;
        cnt += (unsigned long long)__popc(m);
        while (m) {
            const int j = __ffs((int)m) - 1;
            m &= m - 1u;
            acc += kc_mix64_hd(lw.code64(j, kmask));
        }
    });
    block_add_u64(acc, cnt, out);
}

__global__ void __launch_bounds__(256) sparse_fp_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                                        uint64_t n, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0, cnt = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = counts[i];
        acc += c * kc_mix64_hd(keys[i]);
        cnt += c;
        if (i && keys[i] <= keys[i - 1]) atomicAdd(out + 2, 1ull);  // not strictly ascending here (never, in a valid result)
    }
    block_add_u64(acc, cnt, out);
}

__global__ void __launch_bounds__(256) dense_fp_kernel(const uint32_t* __restrict__ table, uint64_t nbins,
                                                       unsigned long long* __restrict__ out) {
    unsigned long long acc = 0, cnt = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = table[i];
        if (c) {
            acc += c * kc_mix64_hd(i);
            cnt += c;
        }
    }
    block_add_u64(acc, cnt, out);
}

int fetch(kc_ctx* ctx, unsigned long long* d_out, uint64_t* h_fp, uint64_t* h_total, uint64_t* h_third = nullptr) {
    unsigned long long h[3] = {0, 0, 0};
    KC_CUDA(ctx, cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_fp) *h_fp = h[0];
    if (h_total) *h_total = h[1];
    if (h_third) *h_third = h[2];
    return KC_OK;
}

}  // namespace

extern "C" {

int kc_window_fingerprint(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, uint64_t* h_fp, uint64_t* h_windows) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_K) return kc_set_error(ctx, KC_ERR_INVALID, "k must be 1..%d, got %d", KC_MAX_K, k);
    if (!d_data && nbytes) return kc_set_error(ctx, KC_ERR_INVALID, "null data");
    DeviceGuard dg(ctx->device);
    int rc = kc_scratch2_reserve(ctx, 256);
    if (rc) return rc;
    unsigned long long* d_out = (unsigned long long*)ctx->scratch2;
    KC_CUDA(ctx, cudaMemsetAsync(d_out, 0, 24, ctx->stream));
    const uint64_t nwin = nbytes >= (uint64_t)k ? nbytes - k + 1 : 0;
    if (nwin) {
        const ScanGeom g = kc_make_geom(d_data, nbytes, 0, nwin, k);
        const uint64_t ngroups = g.g_end - g.g_begin;
        uint64_t want = (ngroups + 63) / 64;
        const uint64_t maxg = (uint64_t)ctx->sm_count * 8;
        const int grid = (int)(want < 1 ? 1 : (want > maxg ? maxg : want));
        if (k <= 17)
            KC_LAUNCH(window_fp_kernel<1>, grid, 256, 0, ctx->stream, g, d_out);
        else
            KC_LAUNCH(window_fp_kernel<2>, grid, 256, 0, ctx->stream, g, d_out);
        KC_LAUNCH_CHECK(ctx, "window_fp_kernel");
    }
    return fetch(ctx, d_out, h_fp, h_windows);
}

int kc_sparse_fingerprint(kc_ctx* ctx, const kc_sparse* s, uint64_t* h_fp, uint64_t* h_total, uint64_t* h_descents) {
    if (!ctx || !s) return KC_ERR_INVALID;
    DeviceGuard dg(ctx->device);
    int rc = kc_scratch2_reserve(ctx, 256);
    if (rc) return rc;
    unsigned long long* d_out = (unsigned long long*)ctx->scratch2;
    KC_CUDA(ctx, cudaMemsetAsync(d_out, 0, 24, ctx->stream));
    if (s->size) {
        KC_LAUNCH(sparse_fp_kernel, ctx->sm_count * 8, 256, 0, ctx->stream, s->d_keys, s->d_counts, s->size, d_out);
        KC_LAUNCH_CHECK(ctx, "sparse_fp_kernel");
    }
    return fetch(ctx, d_out, h_fp, h_total, h_descents);
}

int kc_dense_fingerprint(kc_ctx* ctx, const uint32_t* d_table, int k, uint64_t* h_fp, uint64_t* h_total) {
    if (!ctx || !d_table) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    DeviceGuard dg(ctx->device);
    int rc = kc_scratch2_reserve(ctx, 256);
    if (rc) return rc;
    unsigned long long* d_out = (unsigned long long*)ctx->scratch2;
    KC_CUDA(ctx, cudaMemsetAsync(d_out, 0, 24, ctx->stream));
    KC_LAUNCH(dense_fp_kernel, ctx->sm_count * 8, 256, 0, ctx->stream, d_table, (uint64_t)1 << (2 * k), d_out);
    KC_LAUNCH_CHECK(ctx, "dense_fp_kernel");
    return fetch(ctx, d_out, h_fp, h_total);
}

}  // extern "C"

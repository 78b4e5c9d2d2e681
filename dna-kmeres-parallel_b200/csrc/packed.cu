// packed.cu — "next" row f4: the 2-bit packed sequence store the reference sketches in comments
// (main.cu:78-86, utils.h:65-92: "AACG -> 00000110", four bases per byte, the FIRST base in the two
// most significant bits, A=00 C=01 G=10 T=11), plus what the sketch lacks: a validity bitmap, because
// real sequences hold N, separators and lower case and the count semantics (a window counts iff all
// its bytes are upper-case ACGT, main.cu:643-644) must survive packing.
//
//   packed  : (n+3)/4 bytes, byte i/4 holds base i at bits [6 - 2(i%4), 8 - 2(i%4));  invalid -> 00
//   badmask : one bit per base, bit i%32 of 32-bit word i/32 (LSB first); set = byte i was not ACGT
//   => 0.25 + 0.125 bytes per base at rest instead of 1.
//
// Counting from the store (kc_count_dense_packed) unpacks 2^30-base chunks into an ASCII scratch and
// runs the ordinary dense path on each (windows that START in the chunk; the chunk carries a (k-1)-base
// halo), so every verified kernel is reused and the result equals kc_count_dense of the original
// bytes.  The counting kernels are bound by shared-memory atomics and decode, not by HBM, so a kernel
// that scans the packed words directly would save little; the store is about HBM capacity (a 30 Gbp
// read set is 11 GB packed).
#include "common.cuh"

namespace {

// 16 bases, 2 bits each, base j at bits [2j, 2j+2)  ->  four bytes of the store (and back: the
// permutation is its own inverse): within every byte the four 2-bit groups change places 0<->3, 1<->2
__device__ __forceinline__ uint32_t swap_groups(uint32_t x) {
    return ((x >> 6) & 0x03030303u) | ((x >> 2) & 0x0C0C0C0Cu) | ((x << 2) & 0x30303030u) | ((x << 6) & 0xC0C0C0C0u);
}

__global__ void __launch_bounds__(256)
pack_kernel(const char* __restrict__ data, uint64_t n, uint8_t* __restrict__ packed, uint32_t* __restrict__ badmask) {
    const uint64_t nblk = (n + 15) >> 4;                       // 16-base blocks
    const uint64_t nblk_r = (nblk + 31) & ~(uint64_t)31;       // whole warps: the mask words need both lanes of a pair
    const int lane = threadIdx.x & 31;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk_r; b += (uint64_t)gridDim.x * blockDim.x) {
        Decoded16 d;
        d.packed = 0;
        d.bad = 0xFFFFu;
        const uint64_t p0 = b << 4;
        if (p0 + 16 <= n) {
            d = kc_decode16(kc_ldg_stream(reinterpret_cast<const uint4*>(data) + b));
        } else if (p0 < n) {  // the last, partial block: byte by byte
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint64_t i = p0; i < n; i++) w[(i - p0) >> 2] |= (uint32_t)(uint8_t)data[i] << (8 * ((i - p0) & 3));
            d = kc_decode16(make_uint4(w[0], w[1], w[2], w[3]));
            d.bad |= 0xFFFFu << (n - p0);
            d.bad &= 0xFFFFu;
        }
        const uint32_t code = d.packed & ~kc_spread_bad(d.bad);  // invalid bases pack as 00
        if (p0 < n) {
            const uint32_t out = swap_groups(code);
            const uint64_t nb = (n - p0 + 3) >> 2;  // bytes of this block that exist (1..4)
            if (nb >= 4)
                reinterpret_cast<uint32_t*>(packed)[b] = out;
            else
                for (uint64_t q = 0; q < nb; q++) packed[(b << 2) + q] = (uint8_t)(out >> (8 * q));
        }
        // mask word = this lane's 16 bits (even lane) + the next lane's 16 bits
        const uint32_t other = __shfl_down_sync(0xffffffffu, d.bad, 1);
        if (!(lane & 1) && (b << 4) < n) badmask[b >> 1] = (d.bad & 0xFFFFu) | (other << 16);
    }
}

__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t* __restrict__ packed, const uint32_t* __restrict__ badmask, uint64_t first, uint64_t n,
              char* __restrict__ out) {
    // bases [first, first + n) of the store -> out[0, n);  `first` is a multiple of 32 (host)
    const uint64_t nblk = (n + 15) >> 4;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (first >> 4) + b;  // 16-base block of the store
        const uint64_t p0 = b << 4;
        uint32_t w = 0;
        const uint8_t* src = packed + (g << 2);
        const uint64_t nbytes = (n - p0 >= 16) ? 4 : ((n - p0 + 3) >> 2);
        if (nbytes == 4 && ((uintptr_t)src & 3) == 0)
            w = *reinterpret_cast<const uint32_t*>(src);
        else
            for (uint64_t q = 0; q < nbytes; q++) w |= (uint32_t)src[q] << (8 * q);
        const uint32_t code = swap_groups(w);  // base j at bits [2j, 2j+2)
        const uint32_t bad = (badmask[g >> 1] >> (16 * (g & 1))) & 0xFFFFu;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t x = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int j = 4 * q + t;
                const uint32_t c = (code >> (2 * j)) & 3u;
                const uint32_t ch = ((bad >> j) & 1u) ? 0x4Eu : ((0x54474341u >> (8 * c)) & 0xFFu);  // 'N' or "ACGT"[c]
                x |= ch << (8 * t);
            }
            o[q] = x;
        }
        if (n - p0 >= 16 && ((uintptr_t)out & 15) == 0) {
            reinterpret_cast<uint4*>(out)[b] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (uint64_t i = 0; i < 16 && p0 + i < n; i++) out[p0 + i] = (char)(o[i >> 2] >> (8 * (i & 3)));
        }
    }
}

}  // namespace

extern "C" {

uint64_t kc_packed_bytes(uint64_t nbases) { return (nbases + 3) / 4; }
uint64_t kc_badmask_bytes(uint64_t nbases) { return (nbases + 31) / 32 * 4; }

int kc_pack_2bit(kc_ctx* ctx, const char* d_data, uint64_t nbytes, void* d_packed, uint32_t* d_badmask, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (nbytes == 0) return KC_OK;
    if (!d_data || !d_packed || !d_badmask) return kc_set_error(ctx, KC_ERR_INVALID, "kc_pack_2bit: null pointer");
    if (((uintptr_t)d_data & 15) || ((uintptr_t)d_packed & 3))
        return kc_set_error(ctx, KC_ERR_INVALID, "kc_pack_2bit: d_data must be 16-byte and d_packed 4-byte aligned");
    DeviceGuard dg(ctx->device);
    const uint64_t nblk = (nbytes + 15) >> 4;
    const uint64_t want = (nblk + 255) / 256;
    const int grid = (int)(want > (uint64_t)ctx->sm_count * 8 ? (uint64_t)ctx->sm_count * 8 : want);
    KC_LAUNCH(pack_kernel, grid, 256, 0, (cudaStream_t)stream, d_data, nbytes, (uint8_t*)d_packed, d_badmask);
    KC_LAUNCH_CHECK(ctx, "pack_kernel");
    return KC_OK;
}

static int unpack_range(kc_ctx* ctx, const void* d_packed, const uint32_t* d_badmask, uint64_t first, uint64_t n, char* d_out,
                        cudaStream_t st) {
    const uint64_t nblk = (n + 15) >> 4;
    const uint64_t want = (nblk + 255) / 256;
    const int grid = (int)(want > (uint64_t)ctx->sm_count * 8 ? (uint64_t)ctx->sm_count * 8 : (want ? want : 1));
    KC_LAUNCH(unpack_kernel, grid, 256, 0, st, (const uint8_t*)d_packed, d_badmask, first, n, d_out);
    KC_LAUNCH_CHECK(ctx, "unpack_kernel");
    return KC_OK;
}

int kc_unpack_2bit(kc_ctx* ctx, const void* d_packed, const uint32_t* d_badmask, uint64_t nbases, char* d_data_out, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (nbases == 0) return KC_OK;
    if (!d_packed || !d_badmask || !d_data_out) return kc_set_error(ctx, KC_ERR_INVALID, "kc_unpack_2bit: null pointer");
    DeviceGuard dg(ctx->device);
    return unpack_range(ctx, d_packed, d_badmask, 0, nbases, d_data_out, (cudaStream_t)stream);
}

int kc_count_dense_packed(kc_ctx* ctx, const void* d_packed, const uint32_t* d_badmask, uint64_t nbases, int k, uint32_t* d_table) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!d_table || (nbases && (!d_packed || !d_badmask))) return kc_set_error(ctx, KC_ERR_INVALID, "kc_count_dense_packed: null pointer");
    DeviceGuard dg(ctx->device);
    cudaStream_t st = ctx->stream;
    KC_CUDA(ctx, cudaMemsetAsync(d_table, 0, sizeof(uint32_t) << (2 * k), st));
    if (nbases >= (uint64_t)k) {
        const uint64_t nwin = nbases - k + 1;
        static const uint64_t chunk_env = getenv("KC_PACKED_CHUNK") ? strtoull(getenv("KC_PACKED_CHUNK"), nullptr, 0) : 0;  // test aid
        // window starts per chunk, a multiple of 32.  2^30: the partition path's fixed cost per call (~0.1 ms: 2048
        // sub-table flushes) is paid once per Gbp, and 1 GiB of scratch is nothing on a 180 GB part
        const uint64_t chunk = chunk_env ? (chunk_env + 31) / 32 * 32 : (1ull << 30);
        const uint64_t need = (nwin < chunk ? nwin : chunk) + KC_MAX_DENSE_K + 64;
        int rc = kc_scratch2_reserve(ctx, (size_t)need);
        if (rc) return rc;
        char* ascii = (char*)ctx->scratch2;
        for (uint64_t w0 = 0; w0 < nwin; w0 += chunk) {
            const uint64_t w1 = (w0 + chunk < nwin) ? w0 + chunk : nwin;
            const uint64_t len = w1 - w0 + k - 1;  // bases [w0, w1 + k - 1): the chunk and its halo
            rc = unpack_range(ctx, d_packed, d_badmask, w0, len, ascii, st);
            if (rc) return rc;
            rc = kc_count_dense_range_async(ctx, ascii, len, 0, w1 - w0, k, d_table, KC_DENSE_AUTO, st);
            if (rc) return rc;
        }
    }
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    return KC_OK;
}

}  // extern "C"

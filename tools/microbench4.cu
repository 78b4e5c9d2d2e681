// microbench4.cu — does TMA staging help the input scan?  (north_star: "TMA staging feeds this path";
// decide by measurement.)
//
// The same consumer — ASCII -> 2-bit decode of every 16-byte block, XOR-reduced so that nothing is optimised away —
// fed two ways over the same 3.1 GB buffer:
//   A  direct:  ld.global.nc.L1::no_allocate.v4 per lane, three blocks in flight per lane (what WarpScanner /
//               part_scatter7v2_kernel do; seven in flight there)
//   B  TMA:     one elected thread issues cp.async.bulk global -> shared (1-D bulk copy, mbarrier complete_tx)
//               of 16 KiB tiles into a ring of STAGES tiles per CTA; the CTA waits for a tile's mbarrier, every
//               thread reads its 4 x 16 bytes with conflict-free ld.shared.v4, decodes, and the stage is handed
//               back behind a CTA barrier.
// Reported: GB/s of each (best of 5).  nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../dna-kmeres-parallel_b200/csrc/common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
int kc_set_error(kc_ctx*, int c, const char*, ...) { return c; }

template <int DEPTH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_direct(const uint4* __restrict__ base, uint64_t nblocks16, uint32_t* out) {
    // every warp owns a contiguous run of 512-byte groups
    const int lane = threadIdx.x & 31;
    const uint64_t ngroups = nblocks16 / 32;
    const uint64_t nwarps = (uint64_t)gridDim.x * (THREADS / 32);
    const uint64_t w = (uint64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t gb = w * ngroups / nwarps, ge = (w + 1) * ngroups / nwarps;
    if (gb >= ge) return;
    const uint4* ptr = base + gb * 32 + lane;
    uint4 raw[DEPTH];
#pragma unroll
    for (int q = 0; q < DEPTH; q++) raw[q] = kc_ldg_stream(ptr + 32 * q);  // reads past the end stay inside the allocation (host)
    uint32_t acc = 0;
    for (uint64_t g = gb; g < ge; g += DEPTH) {
#pragma unroll
        for (int q = 0; q < DEPTH; q++) {
            const Decoded16 d = kc_decode16(raw[q]);
            raw[q] = kc_ldg_stream(ptr + 32 * (q + DEPTH));
            acc ^= d.packed + d.bad;
        }
        ptr += 32 * DEPTH;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

template <int THREADS, int STAGES, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_tma(const char* __restrict__ base, uint64_t ntiles, uint32_t* out) {
    constexpr uint32_t TILE = THREADS * 64;  // bytes: four 16-byte blocks per thread
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[STAGES];
    const uint32_t s_tiles = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_bars = (uint32_t)__cvta_generic_to_shared(bars);
    const int tid = threadIdx.x;
    // tiles of this CTA: a contiguous run
    const uint64_t tb = (uint64_t)blockIdx.x * ntiles / gridDim.x, te = (uint64_t)(blockIdx.x + 1) * ntiles / gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(s_bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && tb + s < te; s++) {
            mbar_expect_tx(s_bars + 8 * s, TILE);
            tma_load_1d(s_tiles + s * TILE, base + (tb + s) * TILE, TILE, s_bars + 8 * s);
        }
    }
    uint32_t acc = 0;
    for (uint64_t t = tb; t < te; t++) {
        const uint32_t s = (uint32_t)((t - tb) % STAGES), parity = (uint32_t)(((t - tb) / STAGES) & 1);
        mbar_wait(s_bars + 8 * s, parity);
        const uint32_t src = s_tiles + s * TILE + tid * 16;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const Decoded16 d = kc_decode16(smem_ld128(src + q * (THREADS * 16)));  // lane-contiguous 16-byte pieces: conflict-free
            acc ^= d.packed + d.bad;
        }
        __syncthreads();  // every thread has read stage s
        if (tid == 0 && t + STAGES < te) {
            mbar_expect_tx(s_bars + 8 * s, TILE);
            tma_load_1d(s_tiles + s * TILE, base + (t + STAGES) * TILE, TILE, s_bars + 8 * s);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
float best_of(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const uint64_t bytes = 3100000000ull / 65536 * 65536;  // whole 64 KiB units
    char* d;
    uint32_t* out;
    CK(cudaMalloc(&d, bytes + (1 << 20)));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(d, 'A', bytes + (1 << 20)));
    printf("device %s, %d SMs; scan of %.2f GB, decode + XOR per 16-byte block\n", prop.name, sms, bytes / 1e9);
    const uint64_t nblocks16 = bytes / 16;
    {
        float ms = best_of([&] { k_direct<3, 1024, 1><<<sms, 1024>>>((const uint4*)d, nblocks16 / (32 * 3) * (32 * 3), out); });
        printf("A direct ld.global.nc.v4, 3 in flight, 1024 thr x 1 CTA/SM   %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        ms = best_of([&] { k_direct<7, 1024, 1><<<sms, 1024>>>((const uint4*)d, nblocks16 / (32 * 7) * (32 * 7), out); });
        printf("A direct ld.global.nc.v4, 7 in flight, 1024 thr x 1 CTA/SM   %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        ms = best_of([&] { k_direct<4, 256, 4><<<sms * 4, 256>>>((const uint4*)d, nblocks16 / (32 * 4) * (32 * 4), out); });
        printf("A direct ld.global.nc.v4, 4 in flight, 256 thr x 4 CTA/SM    %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    {
        constexpr int T = 256, S = 4;
        const size_t smem = (size_t)T * 64 * S;
        CK(cudaFuncSetAttribute(k_tma<T, S, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint64_t ntiles = bytes / (T * 64);
        float ms = best_of([&] { k_tma<T, S, 3><<<sms * 3, T, smem>>>(d, ntiles, out); });
        printf("B TMA cp.async.bulk 16 KiB tiles, %d stages, %d thr x 3 CTA/SM  %.3f ms  %.0f GB/s\n", S, T, ms, bytes / ms / 1e6);
    }
    {
        constexpr int T = 512, S = 3;
        const size_t smem = (size_t)T * 64 * S;
        CK(cudaFuncSetAttribute(k_tma<T, S, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint64_t ntiles = bytes / (T * 64);
        float ms = best_of([&] { k_tma<T, S, 2><<<sms * 2, T, smem>>>(d, ntiles, out); });
        printf("B TMA cp.async.bulk 32 KiB tiles, %d stages, %d thr x 2 CTA/SM  %.3f ms  %.0f GB/s\n", S, T, ms, bytes / ms / 1e6);
    }
    {
        constexpr int T = 1024, S = 3;
        const size_t smem = (size_t)T * 64 * S;
        CK(cudaFuncSetAttribute(k_tma<T, S, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint64_t ntiles = bytes / (T * 64);
        float ms = best_of([&] { k_tma<T, S, 1><<<sms, T, smem>>>(d, ntiles, out); });
        printf("B TMA cp.async.bulk 64 KiB tiles, %d stages, %d thr x 1 CTA/SM  %.3f ms  %.0f GB/s\n", S, T, ms, bytes / ms / 1e6);
    }
    return 0;
}

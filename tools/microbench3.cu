// microbench3.cu — the primitive rates the round-2 kernel variants are argued from, measured with as
// little ALU around them as possible (microbench.cu spends 5 instructions per operation):
//   1. shared-memory increments: red.shared.add / atom.shared.add, random vs conflict-free vs one bank,
//      full warps vs half-active warps (the pass-2 count kernels and their paired/trio variants)
//   2. plain st.shared / ld.shared with the same address patterns (bank-conflict reference)
//   3. bin flush: a few lanes per warp move 64 bytes shared -> global with 4 x (LDS.128 + STG.128),
//      with 2 x STG.256, or with one TMA bulk copy (the scatter kernel's flush variants)
// Build: make -C tools microbench3 ; run: tools/microbench3     (needs a B200)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum { PAT_RANDOM = 0, PAT_LANE_BANK = 1, PAT_ONE_BANK = 2 };
enum { OP_RED = 0, OP_ATOM = 1, OP_STS = 2, OP_LDS = 3 };

// every thread runs `iters` x 8 operations on a 32768-word table; the 8 word offsets live in
// registers and advance by a per-pattern stride, so an operation costs one add + the memory op
template <int OP, int PAT, int HALF>
__global__ void __launch_bounds__(1024, 1) k_smem(int iters, uint32_t* out) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t off[8], sink = 0;
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
#pragma unroll
    for (int u = 0; u < 8; u++) {
        s = s * 1664525u + 1013904223u;
        const uint32_t w = (s >> 12) & 32767u;
        off[u] = PAT == PAT_RANDOM ? w : PAT == PAT_LANE_BANK ? ((w & ~31u) | lane) : (w & ~31u);
    }
    // stride keeps the pattern: random -> an odd word stride per thread; lane/one bank -> multiples of 32 words
    const uint32_t stride = PAT == PAT_RANDOM ? (((s >> 9) | 1u) & 32767u) : 32u * (1u + ((s >> 20) & 63u));
    if (!HALF || (lane & 1)) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t a = base + off[u] * 4;
                if (OP == OP_RED) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(1u) : "memory");
                if (OP == OP_ATOM) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(1u) : "memory"); sink += r; }
                if (OP == OP_STS) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(sink) : "memory");
                if (OP == OP_LDS) { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); sink += r; }
                off[u] = (off[u] + stride) & 32767u;
            }
        }
    }
    __syncthreads();
    uint32_t acc = sink;
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) acc += tab[i];
    if (acc == 0x12345678u) out[0] = acc;
}

// flush: in every warp the lanes with (lane % 16 == 0) move one 64-byte bin per iteration
template <int MODE>  // 0: 4 x (LDS.128 + STG.128), 1: 4 x LDS.128 + 2 x STG.256, 2: one bulk copy
__global__ void __launch_bounds__(1024, 1) k_flush(int iters, uint32_t* __restrict__ dst, uint32_t* out) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 777u;
    uint32_t* my = dst + ((size_t)blockIdx.x * 1024 + threadIdx.x) * 16 * 64;  // 64 chunks of 64 bytes per thread
    if ((lane & 15) == 0) {
        for (int it = 0; it < iters; it++) {
            s = s * 1664525u + 1013904223u;
            const uint32_t bin = (s >> 12) & 2047u;  // 2048 bins of 64 bytes
            const uint32_t src = base + bin * 64;
            uint32_t* g = my + (it & 63) * 16;
            if (MODE == 2) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 64;" ::"l"(g), "r"(src) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            } else {
                uint4 v[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[q].x), "=r"(v[q].y), "=r"(v[q].z), "=r"(v[q].w) : "r"(src + 16 * q) : "memory");
                if (MODE == 0) {
#pragma unroll
                    for (int q = 0; q < 4; q++) reinterpret_cast<uint4*>(g)[q] = v[q];
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q += 2)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(g + 4 * q), "r"(v[q].x), "r"(v[q].y), "r"(v[q].z),
                                     "r"(v[q].w), "r"(v[q + 1].x), "r"(v[q + 1].y), "r"(v[q + 1].z), "r"(v[q + 1].w)
                                     : "memory");
                }
            }
        }
    }
    if (s == 0x12345678u) out[0] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

template <int OP, int PAT, int HALF>
static void run_smem(const char* name, int sms, double ghz, uint32_t* d_out) {
    const int iters = 4096;
    auto kern = k_smem<OP, PAT, HALF>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    const float ms = time_ms([&] { kern<<<sms, 1024, 131072>>>(iters, d_out); });
    const double lane_ops = (double)sms * 1024 * iters * 8 / (HALF ? 2 : 1);
    printf("%-46s %8.3f ms  %7.2f lane-ops/clk/SM  (%.2f clk per warp instruction)\n", name, ms, lane_ops / (ms * 1e-3) / sms / (ghz * 1e9),
           (ms * 1e-3) * ghz * 1e9 / ((double)32 * iters * 8));
}

template <int MODE>
static void run_flush(const char* name, int sms, double ghz, uint32_t* d_dst, uint32_t* d_out) {
    const int iters = 2048;
    auto kern = k_flush<MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    const float ms = time_ms([&] { kern<<<sms, 1024, 131072>>>(iters, d_dst, d_out); });
    const double flushes = (double)sms * 64 * iters;  // 2 lanes of each of the 32 warps
    printf("%-46s %8.3f ms  %7.1f clk per flush and SM  (%.2f G flushes/s chip-wide)\n", name, ms, (ms * 1e-3) * ghz * 1e9 / (64.0 * iters),
           flushes / (ms * 1e-3) / 1e9);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    printf("device %s, %d SMs, %.3f GHz (clk figures assume this clock)\n", p.name, sms, ghz);
    uint32_t *d_out, *d_dst;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMalloc(&d_dst, (size_t)sms * 1024 * 16 * 64 * 4));
    run_smem<OP_RED, PAT_RANDOM, 0>("red.shared.add  random words", sms, ghz, d_out);
    run_smem<OP_RED, PAT_LANE_BANK, 0>("red.shared.add  lane = bank (conflict-free)", sms, ghz, d_out);
    run_smem<OP_RED, PAT_ONE_BANK, 0>("red.shared.add  all lanes one bank", sms, ghz, d_out);
    run_smem<OP_RED, PAT_RANDOM, 1>("red.shared.add  random, odd lanes only", sms, ghz, d_out);
    run_smem<OP_ATOM, PAT_RANDOM, 0>("atom.shared.add random words (returning)", sms, ghz, d_out);
    run_smem<OP_ATOM, PAT_LANE_BANK, 0>("atom.shared.add lane = bank (returning)", sms, ghz, d_out);
    run_smem<OP_STS, PAT_RANDOM, 0>("st.shared       random words", sms, ghz, d_out);
    run_smem<OP_STS, PAT_LANE_BANK, 0>("st.shared       lane = bank", sms, ghz, d_out);
    run_smem<OP_LDS, PAT_RANDOM, 0>("ld.shared       random words", sms, ghz, d_out);
    run_smem<OP_LDS, PAT_LANE_BANK, 0>("ld.shared       lane = bank", sms, ghz, d_out);
    run_flush<0>("flush 64 B: 4 x LDS.128 + 4 x STG.128 per lane", sms, ghz, d_dst, d_out);
    run_flush<1>("flush 64 B: 4 x LDS.128 + 2 x STG.256 per lane", sms, ghz, d_dst, d_out);
    run_flush<2>("flush 64 B: one TMA bulk copy per lane", sms, ghz, d_dst, d_out);
    return 0;
}

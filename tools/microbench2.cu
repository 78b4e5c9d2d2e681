// microbench2.cu — isolates the streaming scan (load + 2-bit decode + halo shuffles)
// of the partition scatter kernel: prefetch depth x CTA size, no staging, no stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../dna-kmeres-parallel_b200/csrc/common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
int kc_set_error(kc_ctx*, int c, const char*, ...) { return c; }

template <int DEPTH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_scan(const uint4* __restrict__ base, uint64_t ngroups, uint32_t* out) {
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (uint64_t)gridDim.x * (THREADS / 32);
    const uint64_t gpw = (ngroups + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t gb = min(w * gpw, ngroups), ge = min((w + 1) * gpw, ngroups);
    if (gb >= ge) return;
    const uint32_t n = (uint32_t)(ge - gb);
    const uint4* ptr = base + gb * 32 + lane;
    uint4 raw[DEPTH];
#pragma unroll
    for (int q = 0; q < DEPTH; q++) raw[q] = kc_ldg_stream(ptr + 32 * (q + 1));  // reads past the end are inside the allocation
    Decoded16 cur = kc_decode16(kc_ldg_stream(ptr));
    uint32_t acc = 0;
    uint32_t i = 0;
    for (; i + DEPTH <= n; i += DEPTH) {
#pragma unroll
        for (int q = 0; q < DEPTH; q++) {
            const Decoded16 nxt = kc_decode16(raw[q]);
            raw[q] = kc_ldg_stream(ptr + 32 * (DEPTH + 1));
            uint32_t p1 = __shfl_down_sync(0xffffffffu, cur.packed, 1);
            uint32_t b1 = __shfl_down_sync(0xffffffffu, cur.bad, 1);
            const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
            const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
            if (lane == 31) { p1 = n0p; b1 = n0b; }
            const uint32_t B32 = cur.bad | (b1 << 16);
            uint32_t ok = (1u << 21) - 1u;
            if (B32) ok = ~(uint32_t)kc_window_bad((uint64_t)B32 | (0xFFFFull << 32), 12);
#pragma unroll
            for (int t = 0; t < 4; t++) acc ^= __funnelshift_r(cur.packed, p1, 10 * t) & ok;
            cur = nxt;
            ptr += 32;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int DEPTH, int THREADS, int MINB>
void run(const char* name, const void* p, uint64_t bytes, uint32_t* out, int sms, int smem = 0) {
    CK(cudaFuncSetAttribute(k_scan<DEPTH, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const uint64_t ngroups = bytes / 512 - 8;
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_scan<DEPTH, THREADS, MINB><<<sms * MINB, THREADS, smem>>>((const uint4*)p, ngroups, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("scan %-28s smem=%3dKB: %.3f ms  %.1f GB/s\n", name, smem / 1024, ms, bytes / ms / 1e6);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const uint64_t bytes = 3100000000ull / 512 * 512;
    void* p;
    CK(cudaMalloc(&p, bytes + 65536));
    CK(cudaMemset(p, 'A', bytes + 65536));
    {   // incompressible content: every 8th byte becomes a pseudo-random letter
        char* h = (char*)malloc(1 << 24);
        uint32_t x = 12345;
        for (int i = 0; i < (1 << 24); i++) { x = x * 1664525u + 1013904223u; h[i] = "ACGT"[x >> 30]; }
        for (uint64_t off = 0; off < bytes; off += (1 << 24)) {
            uint64_t n = bytes - off < (1 << 24) ? bytes - off : (1 << 24);
            CK(cudaMemcpy((char*)p + off, h, n, cudaMemcpyHostToDevice));
        }
        free(h);
    }
    uint32_t* out;
    CK(cudaMalloc(&out, 64));
    for (int smem : {0, 100 * 1024, 144 * 1024, 176 * 1024, 208 * 1024}) {
        run<2, 1024, 1>("depth2 1024thr x1", p, bytes, out, sms, smem);
        run<3, 1024, 1>("depth3 1024thr x1", p, bytes, out, sms, smem);
        run<4, 1024, 1>("depth4 1024thr x1", p, bytes, out, sms, smem);
    }
    run<4, 512, 1>("depth4 512thr x1", p, bytes, out, sms, 208 * 1024);
    run<6, 512, 1>("depth6 512thr x1", p, bytes, out, sms, 208 * 1024);
    return 0;
}

// microbench.cu — measures the primitives the dense k-mer kernels are bounded by
// on this B200: streaming 128-bit loads, shared-memory atomics on random bins,
// global REDs on a 64 MiB table.  Build: make -C tools ; run: tools/microbench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// LCG: one IMAD per draw, high bits are the good ones
__device__ __forceinline__ uint32_t xs(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 12; }

__global__ void __launch_bounds__(1024, 1) k_stream(const uint4* __restrict__ p, uint64_t n16, uint32_t* out) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// every thread does `iters` x UNROLL atomics on random bins of a `words`-word smem table
template <int RET>
__global__ void __launch_bounds__(1024, 1) k_smem_atomic(int words, int iters, uint32_t* out) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    uint32_t sink = 0;
    const uint32_t mask = words - 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t r = xs(s);
            if (RET) sink += atomicAdd(&tab[r & mask], 1u);
            else atomicAdd(&tab[r & mask], 1u);
        }
    }
    __syncthreads();
    uint32_t a = sink;
    for (int i = threadIdx.x; i < words; i += blockDim.x) a += tab[i];
    if (a == 0x12345678u) out[0] = a;
}

// conflict-free variant: lane l always hits bank l
__global__ void __launch_bounds__(1024, 1) k_smem_atomic_nc(int words, int iters, uint32_t* out) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t mask = words - 1;
    const uint32_t lane = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t r = xs(s);
            atomicAdd(&tab[((r & mask) & ~31u) | lane], 1u);
        }
    }
    __syncthreads();
    uint32_t a = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) a += tab[i];
    if (a == 0x12345678u) out[0] = a;
}

__global__ void __launch_bounds__(1024, 1) k_smem_store(int words, int iters, uint32_t* out) {
    extern __shared__ uint32_t tab[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * 1024 + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t mask = words - 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t r = xs(s);
            tab[r & mask] = r;
        }
    }
    __syncthreads();
    uint32_t a = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) a += tab[i];
    if (a == 0x12345678u) out[0] = a;
}

__global__ void __launch_bounds__(256) k_global_red(uint32_t* table, uint32_t mask, int iters) {
    uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 999u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) atomicAdd(&table[xs(s) & mask], 1u);
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    uint32_t* out;
    CK(cudaMalloc(&out, 64));
    float ms;

    {   // streaming read of 3 GiB
        const uint64_t bytes = 3ull << 30;
        void* p;
        CK(cudaMalloc(&p, bytes));
        CK(cudaMemset(p, 1, bytes));
        for (int grid_mul : {1, 2}) {
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0));
                k_stream<<<sms * grid_mul, 1024>>>((const uint4*)p, bytes / 16, out);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("stream read 3 GiB grid=%dxSM x1024: %.3f ms  %.1f GB/s\n", grid_mul, ms, bytes / ms / 1e6);
        }
        CK(cudaFree(p));
    }
    {   // shared-memory atomics
        CK(cudaFuncSetAttribute(k_smem_atomic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_smem_atomic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_smem_atomic_nc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_smem_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const int iters = 2000;
        const double ops = (double)sms * 1024 * iters * 8;
        for (int words : {64, 1024, 8192, 32768}) {
            for (int variant = 0; variant < 4; variant++) {
                for (int rep = 0; rep < 2; rep++) {
                    CK(cudaEventRecord(e0));
                    if (variant == 0) k_smem_atomic<0><<<sms, 1024, words * 4>>>(words, iters, out);
                    if (variant == 1) k_smem_atomic<1><<<sms, 1024, words * 4>>>(words, iters, out);
                    if (variant == 2) k_smem_atomic_nc<<<sms, 1024, words * 4>>>(words, iters, out);
                    if (variant == 3) k_smem_store<<<sms, 1024, words * 4>>>(words, iters, out);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                }
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const char* names[] = {"atomicAdd noret random", "atomicAdd ret random", "atomicAdd noret conflict-free", "plain store random"};
                printf("smem %-30s words=%6d: %.3f ms  %.1f Gops/s  (%.2f ops/SM/ns)\n", names[variant], words, ms, ops / ms / 1e6, ops / ms / 1e6 / sms);
            }
        }
    }
    {   // global REDs on random bins
        for (uint32_t mbytes : {1u, 64u, 1024u}) {
            const uint64_t bytes = (uint64_t)mbytes << 20;
            uint32_t* t;
            CK(cudaMalloc(&t, bytes));
            CK(cudaMemset(t, 0, bytes));
            const int iters = 200;
            const double ops = (double)sms * 8 * 256 * iters * 8;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                k_global_red<<<sms * 8, 256>>>(t, (uint32_t)(bytes / 4 - 1), iters);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("global RED random table=%4u MiB: %.3f ms  %.1f Gops/s\n", mbytes, ms, ops / ms / 1e6);
            CK(cudaFree(t));
        }
    }
    return 0;
}

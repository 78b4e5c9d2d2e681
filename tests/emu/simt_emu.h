// simt_emu.h — TEST INFRASTRUCTURE ONLY.  A small SIMT emulator that lets the kernels of
// dna-kmeres-parallel_b200/csrc/*.cu run on the CPU, so their LOGIC (indexing, staging
// protocols, edge handling) can be checked against the oracle in a container without a
// GPU.  It is never linked into libkmerb200.so; the product has no CPU path.
//
//   g++ -x c++ -std=c++17 -DKC_EMU -include tests/emu/simt_emu.h ... csrc/*.cu
//
// Model
//   * every CUDA thread of a CTA is a fiber (own stack, hand-written x86-64 switch);
//     CTAs of a grid run one after the other (legal: no kernel here waits on another
//     CTA);
//   * fibers are switched only inside emulator calls: warp collectives
//     (__shfl*_sync, __ballot_sync, __syncwarp), __syncthreads, and — with
//     KC_EMU_SEED != 0 — at random inside the shared-memory/atomic helpers, so that the
//     barrier-free staging protocols see adversarial interleavings of their threads
//     (sequentially consistent ones: the GPU's weaker ordering is NOT modelled);
//   * dynamic shared memory is filled with 0xCD at launch (reads of uninitialised
//     shared memory show up), device allocations end at a PROT_NONE guard page
//     (reads/writes past the end fault);
//   * the CUDA runtime calls the host code uses are synchronous stand-ins.
#pragma once
#ifndef KC_EMU
#error "simt_emu.h is for -DKC_EMU builds only"
#endif

#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/mman.h>

#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#define __CUDACC__ 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct uint4 {
    uint32_t x, y, z, w;
} __attribute__((aligned(16)));
struct uint2 {
    uint32_t x, y;
} __attribute__((aligned(8)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// CUDA's mixed-type min/max
template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a > (T)b ? (T)a : (T)b;
}

namespace emu {

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    int state = 0;  // 0 runnable, 1 blocked, 2 done
    int run_pos = -1;
};
struct Warp {
    uint32_t arrived = 0;
    uint32_t gen = 0;
    uint32_t vals[2][32];
    uint32_t ballot[2] = {0, 0};
};

struct State {
    dim3 grid, block;
    size_t smem_bytes = 0;
    char* dyn_smem = nullptr;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<int> runnable;
    int cur = -1;
    int live = 0;
    int bar_arrived = 0;
    std::vector<int> bar_waiters;
    void* sched_sp = nullptr;
    const std::function<void()>* body = nullptr;
    uint64_t rng = 0;
    uint32_t preempt_mask = 0;  // yield when (rng & mask) == 0; 0 = never
    uint64_t launches = 0, switches = 0;
    bool in_kernel = false;
};
extern State S;
extern uint64_t stats[16];  // KC_STAT(i) counters of rare paths, printed at exit when KC_EMU_STATS is set
extern dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void yield_blocked();      // current fiber blocks until somebody makes it runnable
void maybe_preempt();      // random voluntary yield (stays runnable)
void maybe_preempt_always();  // unconditional yield (spin loops must let the others run)
void make_runnable(int f);
uint32_t warp_exchange(uint32_t mask, uint32_t v, int kind, int arg);  // shfl/ballot/syncwarp
void syncthreads();
[[noreturn]] void fail(const char* fmt, ...);

static constexpr uint32_t SMEM_WINDOW = 0x400;  // shared-window address of dynamic smem byte 0
static inline char* smem_ptr(uint32_t saddr, uint32_t nbytes) {
    if (saddr < SMEM_WINDOW || (size_t)(saddr - SMEM_WINDOW) + nbytes > S.smem_bytes || (saddr & (nbytes - 1)))
        fail("shared-memory access out of bounds or misaligned: addr 0x%x size %u (dynamic smem %zu bytes)", saddr, nbytes,
             S.smem_bytes);
    return S.dyn_smem + (saddr - SMEM_WINDOW);
}

}  // namespace emu

#define threadIdx (::emu::threadIdx_)
#define blockIdx (::emu::blockIdx_)
#define blockDim (::emu::blockDim_)
#define gridDim (::emu::gridDim_)

// ---- device intrinsics --------------------------------------------------------------
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (sh & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (sh & 31)) >> 32);
}
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}

enum { EMU_SHFL_IDX = 0, EMU_SHFL_DOWN = 1, EMU_SHFL_UP = 2, EMU_SHFL_XOR = 3, EMU_BALLOT = 4, EMU_SYNCWARP = 5, EMU_REDUCE_ADD = 6, EMU_MATCH_ANY = 7 };
template <class T>
static inline T emu_shfl(uint32_t mask, T v, int kind, int arg) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "shuffle of 32/64-bit values");
    if (sizeof(T) == 4) {
        uint32_t u;
        memcpy(&u, &v, 4);
        u = emu::warp_exchange(mask, u, kind, arg);
        memcpy(&v, &u, 4);
        return v;
    }
    uint64_t u;
    memcpy(&u, &v, 8);
    const uint32_t lo = emu::warp_exchange(mask, (uint32_t)u, kind, arg);
    const uint32_t hi = emu::warp_exchange(mask, (uint32_t)(u >> 32), kind, arg);
    u = ((uint64_t)hi << 32) | lo;
    memcpy(&v, &u, 8);
    return v;
}
template <class T>
static inline T __shfl_sync(uint32_t mask, T v, int src) { return emu_shfl(mask, v, EMU_SHFL_IDX, src); }
template <class T>
static inline T __shfl_down_sync(uint32_t mask, T v, unsigned d) { return emu_shfl(mask, v, EMU_SHFL_DOWN, (int)d); }
template <class T>
static inline T __shfl_up_sync(uint32_t mask, T v, unsigned d) { return emu_shfl(mask, v, EMU_SHFL_UP, (int)d); }
template <class T>
static inline T __shfl_xor_sync(uint32_t mask, T v, int m) { return emu_shfl(mask, v, EMU_SHFL_XOR, m); }
template <class T>
static inline uint32_t __match_any_sync(uint32_t mask, T v) {  // lanes of `mask` holding the same value as this one
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "match of 32/64-bit values");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    uint32_t m = emu::warp_exchange(mask, (uint32_t)u, EMU_MATCH_ANY, 0);
    if (sizeof(T) == 8) m &= emu::warp_exchange(mask, (uint32_t)(u >> 32), EMU_MATCH_ANY, 0);
    return m;
}
static inline uint32_t __ballot_sync(uint32_t mask, int pred) { return emu::warp_exchange(mask, pred ? 1u : 0u, EMU_BALLOT, 0); }
static inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(uint32_t mask, int pred) { return (__ballot_sync(mask, pred) & mask) == mask; }
static inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) { return emu::warp_exchange(mask, v, EMU_REDUCE_ADD, 0); }
static inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::warp_exchange(mask, 0, EMU_SYNCWARP, 0); }
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __threadfence() { emu::maybe_preempt(); }
static inline void __threadfence_block() { emu::maybe_preempt(); }

template <class T, class U>
static inline T atomicAdd(T* p, U v) {
    emu::maybe_preempt();
    const T old = *p;
    *p = (T)(old + (T)v);
    return old;
}
template <class T, class U>
static inline T atomicOr(T* p, U v) {
    emu::maybe_preempt();
    const T old = *p;
    *p = (T)(old | (T)v);
    return old;
}
template <class T, class U>
static inline T atomicMax(T* p, U v) {
    emu::maybe_preempt();
    const T old = *p;
    if ((T)v > old) *p = (T)v;
    return old;
}
template <class T, class U>
static inline T atomicExch(T* p, U v) {
    emu::maybe_preempt();
    const T old = *p;
    *p = (T)v;
    return old;
}
template <class T>
static inline T atomicCAS(T* p, T cmp, T val) {
    emu::maybe_preempt();
    const T old = *p;
    if (old == cmp) *p = val;
    return old;
}
static inline size_t __cvta_generic_to_shared(const void* p) {
    const char* c = (const char*)p;
    if (c < emu::S.dyn_smem || c > emu::S.dyn_smem + emu::S.smem_bytes) emu::fail("__cvta_generic_to_shared of a non-dynamic-smem pointer");
    return (size_t)(c - emu::S.dyn_smem) + emu::SMEM_WINDOW;
}

#define KC_STAT(i) (::emu::stats[i]++)
#define KC_DYN_SMEM(T, name) T* const name = reinterpret_cast<T*>(::emu::S.dyn_smem)
#define KC_LAUNCH(kern, grid, block, smem, stream, ...) \
    ::emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })

// ---- CUDA runtime stand-ins ------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
struct cudaDeviceProp {
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin;
    char name[64];
};
cudaError_t emu_cuda_malloc(void** p, size_t n);
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t n) {
    return emu_cuda_malloc((void**)p, n);
}
cudaError_t cudaFree(void* p);
// stream-ordered pool allocation (kc_pool_alloc): plain guarded allocations here
typedef struct emu_pool* cudaMemPool_t;
enum { cudaMemPoolAttrReleaseThreshold = 4, cudaMemPoolAttrReservedMemCurrent = 5, cudaMemPoolAttrUsedMemCurrent = 7 };
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return emu_cuda_malloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) {
    *pool = nullptr;
    return cudaSuccess;
}
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, int, void* v) {
    *(uint64_t*)v = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) {
    *s = (cudaStream_t)malloc(8);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    free(s);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = (cudaEvent_t)calloc(1, 8);
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    free(e);
    return cudaSuccess;
}
// an event holds the wall-clock time of its record (everything is synchronous here), so that host code
// that branches on elapsed times (bench.py's per-kernel table) takes the same branches as on a GPU
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    *reinterpret_cast<double*>(e) = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }  // everything is synchronous here
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(*reinterpret_cast<double*>(b) - *reinterpret_cast<double*>(a));
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetDevice(int* d) {
    *d = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int dev);
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) {
    *f = (size_t)4 << 30;
    *t = (size_t)8 << 30;
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// common.cuh — shared host/device pieces of libkmerb200 (sm_100a only).
//
// The device core is `WarpScanner`: a warp streams 512 aligned bytes per step
// with one coalesced 128-bit load per lane, turns ASCII into a 2-bit packed
// word + a per-base "bad" mask entirely in registers, pulls the (k-1)-base halo
// from the neighbouring lanes with shuffles (and from the next step's lanes
// 0/1 for the last lanes, so every byte is loaded and decoded exactly once),
// and hands each lane its 16 window codes.  K-mer semantics follow the
// reference: window = k consecutive bytes, valid iff all are upper-case ACGT
// (main.cu:641-644, kernels.h:133-140); code is little-endian in the string,
// idx = sum code(s[p]) * 4^p (utils.h:30-47, main.cu:134-135).
#pragma once
#ifndef KC_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/kmer_b200.h"

// Kernel launch and dynamic shared memory go through two macros so that the same sources
// also build against the test-only CPU emulator (tests/emu/simt_emu.h, -DKC_EMU), which
// runs the kernels' logic against the oracle where there is no GPU.  `kern` must be a plain
// identifier (take `auto kern = some_kernel<...>;` first when the name has commas).
#ifndef KC_EMU
#define KC_STAT(i) ((void)0)  // rare-path counters exist in the emulator build only
#define KC_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define KC_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

struct kc_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;      // stream of the synchronous entry points
    cudaStream_t copy_stream = nullptr; // H2D staging of the *_host entry points
    uint64_t launches = 0;
    std::string err;
    // scratch, grown on demand and reused across calls
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* scratch2 = nullptr;
    size_t scratch2_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    uint64_t caller_reusable_bytes = 0;  // device memory the caller holds in its own cache and will hand back (kc_ctx_set_reusable_bytes)
    uint64_t last_h2d_bytes = 0;        // bytes the last *_host* entry point sent over PCIe (kc_ctx_last_h2d_bytes)
    // optional per-kernel timing (kc_ctx_set_timing)
    bool timing = false;
    cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
    int timed_kernels = 0;
};

// what importSeqs leaves behind (main.cu:65-70,34-35): ids, the concatenated sequences, their offsets
struct kc_seqset {
    std::vector<std::string> ids;
    char* data = nullptr;           // sequences, each followed by '\0' (malloc'd); fetched lazily when the
                                    // set was parsed on the GPU (kc_seqset_data)
    size_t data_len = 0, data_cap = 0;
    std::vector<int64_t> offsets;   // num_seqs + 1
    ~kc_seqset() { free(data); }
    uint32_t num_seqs = 0;
    // device copies.  `device` is recorded when they are made: kc_seqset_free / kc_seqset_data must not
    // dereference `owner`, which the caller may have destroyed first (kc_ctx_destroy before kc_seqset_free)
    kc_ctx* owner = nullptr;
    int device = -1;
    char* d_data = nullptr;
    int64_t* d_offsets = nullptr;
};

// result of a sparse count: distinct codes ascending + their counts, both in device memory
struct kc_sparse {
    kc_ctx* ctx = nullptr;
    int device = -1;   // recorded at creation: kc_sparse_free does not dereference ctx (it may be gone)
    uint64_t size = 0;
    uint64_t capacity = 0;  // entries the arrays hold; 0 = exactly `size` (a result that rounds append to has more)
    uint64_t* d_keys = nullptr;
    uint32_t* d_counts = nullptr;
};
// sparse_radix.cu: MSD radix partition + shared-memory leaf sort; *failed = 1 means an
// overflow made the result unusable and nothing was produced (the caller recounts)
int kc_sparse_radix(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, kc_sparse** out, int* failed);

// dense.cu: the direct (shared-bin / global-RED) kernels on a window range; dense_wide.cu: the k = 12
// partition path with seven windows per record
int kc_dense_direct_range(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end, int k,
                          uint32_t* d_table, cudaStream_t st);
int kc_dense_partition_wide(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                            uint32_t* d_table, cudaStream_t st);
int kc_dense_host_plain(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* h_table /* may be NULL */);
int kc_dense_partition_wide2(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                             uint32_t* d_table, cudaStream_t st);

int kc_set_error(kc_ctx* ctx, int code, const char* fmt, ...);
// KC_TRACE=1: host wall-clock between named points of a call, on stderr (measurement aid; a point with
// sync = true first waits for the ctx stream, so the interval before it includes the device work)
void kc_trace(kc_ctx* ctx, const char* what, bool sync = false);
// Temporary and result buffers of the sparse paths come from the device's stream-ordered memory pool
// (cudaMallocAsync on the legacy default stream, which is ordered against the ctx's blocking streams like cudaMalloc /
// cudaFree are, but costs microseconds once the pool is warm: config 4 spent ~200 of 676 ms per call in cudaMalloc /
// cudaFree of multi-GB arrays, gpurun_out r02 call 4).  kc_ctx_create raises the pool's release threshold so that freed
// memory stays with the pool; the driver hands it back when another allocation would otherwise fail, and
// kc_ctx_destroy trims it.  kc_pool_free accepts cudaMalloc'ed pointers too.
cudaError_t kc_pool_alloc(void** p, size_t nbytes);
void kc_pool_free(void* p);
size_t kc_pool_idle_bytes(int device);  // reserved by the pool but not in use: available to kc_pool_alloc
int kc_scratch_reserve(kc_ctx* ctx, size_t nbytes);   // ctx->scratch  >= nbytes
int kc_scratch2_reserve(kc_ctx* ctx, size_t nbytes);  // ctx->scratch2 >= nbytes
void kc_scratch_release(kc_ctx* ctx);                 // free both (after stream sync); they grow again on demand

#define KC_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return kc_set_error((ctx), KC_ERR_CUDA, "%s failed: %s (%s:%d)", #call,          \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define KC_LAUNCH_CHECK(ctx, name)                                                           \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return kc_set_error((ctx), KC_ERR_CUDA, "launch of %s failed: %s", name,         \
                                cudaGetErrorString(e__));                                    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------
// hashing shared with the oracle (oracle/kmer_oracle.c: or_sm64 / or_mix64)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t kc_sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t kc_mix64_hd(uint64_t x) {
    x ^= x >> 33;
    x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33;
    x *= 0xC4CEB9FE1A85EC53ull;
    x ^= x >> 33;
    return x;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// ASCII -> 2-bit, four bases per 32-bit word, SIMD-within-a-register.
//   code = ((x>>1) ^ (x>>2)) & 3      A(0x41)->0 C(0x43)->1 G(0x47)->2 T(0x54)->3
//   valid byte <=> ((x ^ (isT * 0x11)) & 0xF9) == 0x41, isT = bit2 & ~bit1
// (exactly the four upper-case letters pass; see DESIGN.md §3.1 for the proof)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void kc_decode4(uint32_t x, uint32_t& packed8, uint32_t& badbytes) {
    const uint32_t t1 = x >> 1, t2 = x >> 2;
    const uint32_t c = (t1 ^ t2) & 0x03030303u;
    packed8 = (c * 0x01041040u) >> 24;  // b0 | b1<<2 | b2<<4 | b3<<6
    const uint32_t u = t2 & ~t1 & 0x01010101u;
    const uint32_t z = (x ^ (u * 0x11u)) & 0xF9F9F9F9u;
    badbytes = z ^ 0x41414141u;  // non-zero byte <=> invalid base
}

// non-zero bytes of w -> 4 bits
__device__ __forceinline__ uint32_t kc_nzbytes4(uint32_t w) {
    // per byte: (w | (w + 0x7f7f7f7f... )) high bit trick
    const uint32_t t = ((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w;  // bit7 set iff byte != 0
    const uint32_t h = (t >> 7) & 0x01010101u;
    return (h * 0x10204080u) >> 28;  // gather bits 0,8,16,24 -> 0..3
}

// 16 bad bits -> 32-bit mask with both bits of every bad base's 2-bit group set
__device__ __forceinline__ uint32_t kc_spread_bad(uint32_t bad16) {
    uint32_t x = bad16 & 0xFFFFu;          // bit j -> bits 2j, 2j+1
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

struct Decoded16 {
    uint32_t packed;  // 16 bases, 2 bits each, base j at bits [2j,2j+2)
    uint32_t bad;     // bit j set <=> base j is not ACGT (16 bits)
};

__device__ __forceinline__ Decoded16 kc_decode16(uint4 v) {
    uint32_t p0, p1, p2, p3, b0, b1, b2, b3;
    kc_decode4(v.x, p0, b0);
    kc_decode4(v.y, p1, b1);
    kc_decode4(v.z, p2, b2);
    kc_decode4(v.w, p3, b3);
    Decoded16 d;
    d.packed = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
    if ((b0 | b1 | b2 | b3) == 0) {
        d.bad = 0;
    } else {
        d.bad = kc_nzbytes4(b0) | (kc_nzbytes4(b1) << 4) | (kc_nzbytes4(b2) << 8) |
                (kc_nzbytes4(b3) << 12);
    }
    return d;
}

// ---- memory primitives as inline PTX (32-bit shared addresses).  Keeps ptxas from wrapping
// atomicAdd in its warp-aggregation sequence and pins the program order the staging
// protocols rely on.  Under KC_EMU (tests/emu) the same operations are plain C++ with a
// possible fiber switch in front of each, so the protocols meet adversarial interleavings.
#ifndef KC_EMU
__device__ __forceinline__ uint4 kc_ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// coherent (L2) 128-bit load: for scratch a kernel wrote earlier in the same launch
__device__ __forceinline__ uint4 kc_ldg_cg(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ uint32_t kc_ld_cg(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint64_t kc_ld_cg(const uint64_t* p) {
    uint64_t r;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t smem_atom_add(uint32_t saddr, uint32_t v) {
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(saddr), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ void smem_red_add(uint32_t saddr, uint32_t v) {  // no return value
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void smem_st(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void smem_st64(uint32_t saddr, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(saddr), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_ld(uint32_t saddr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ uint64_t smem_ld64(uint32_t saddr) {
    uint64_t r;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(saddr) : "memory");
    return r;
}
// 64-bit global store of two words, 8-byte aligned
__device__ __forceinline__ void kc_stg64(uint32_t* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint4 smem_ld128(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}
// plain RED (no compiler warp-aggregation wrapper around it)
__device__ __forceinline__ void global_red_add(uint32_t* p, uint32_t v) {
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// x >> 31 of a value known to be < 2^31, i.e. a zero the compiler cannot see through:
// OR-ing it into a register copy pins the copy behind the producer of x.
__device__ __forceinline__ uint32_t kc_opaque_zero(uint32_t x) {
    uint32_t z;
    asm volatile("shr.u32 %0, %1, 31;" : "=r"(z) : "r"(x) : "memory");
    return z;
}
#else
static inline uint4 kc_ldg_stream(const uint4* p) { return *p; }
static inline uint4 kc_ldg_cg(const uint4* p) { return *p; }
static inline uint32_t kc_ld_cg(const uint32_t* p) { return *p; }
static inline uint64_t kc_ld_cg(const uint64_t* p) { return *p; }
static inline uint32_t smem_atom_add(uint32_t saddr, uint32_t v) {
    emu::maybe_preempt();
    uint32_t* p = (uint32_t*)emu::smem_ptr(saddr, 4);
    const uint32_t old = *p;
    *p = old + v;
    return old;
}
static inline void smem_red_add(uint32_t saddr, uint32_t v) {
    emu::maybe_preempt();
    *(uint32_t*)emu::smem_ptr(saddr, 4) += v;
}
static inline void smem_st(uint32_t saddr, uint32_t v) {
    emu::maybe_preempt();
    *(uint32_t*)emu::smem_ptr(saddr, 4) = v;
}
static inline void smem_st64(uint32_t saddr, uint64_t v) {
    emu::maybe_preempt();
    *(uint64_t*)emu::smem_ptr(saddr, 8) = v;
}
static inline uint32_t smem_ld(uint32_t saddr) {
    emu::maybe_preempt();
    return *(uint32_t*)emu::smem_ptr(saddr, 4);
}
static inline uint64_t smem_ld64(uint32_t saddr) {
    emu::maybe_preempt();
    return *(uint64_t*)emu::smem_ptr(saddr, 8);
}
static inline void kc_stg64(uint32_t* p, uint32_t a, uint32_t b) {
    emu::maybe_preempt();
    if ((uintptr_t)p & 7) emu::fail("kc_stg64: pointer not 8-byte aligned");
    p[0] = a;
    p[1] = b;
}
static inline uint4 smem_ld128(uint32_t saddr) {
    emu::maybe_preempt();
    return *(uint4*)emu::smem_ptr(saddr, 16);
}
static inline void global_red_add(uint32_t* p, uint32_t v) {
    emu::maybe_preempt();
    *p += v;
}
static inline uint32_t kc_opaque_zero(uint32_t x) { return x >> 31; }
#endif

// bit j of result set <=> any of bad bits j .. j+k-1 set (k in 1..32)
__device__ __forceinline__ uint64_t kc_window_bad(uint64_t B, int k) {
    uint64_t W = 0, span = B;  // span = OR of B>>0 .. B>>(2^bit - 1)
    int shift = 0;
#pragma unroll
    for (int bit = 0; bit < 5; bit++) {
        if (k & (1 << bit)) {
            W |= span >> shift;
            shift += 1 << bit;
        }
        span |= span >> (1 << bit);
    }
    return W;
}

// ---------------------------------------------------------------------------
// WarpScanner: geometry of one scan over a byte buffer.
//   data      : user pointer (any alignment); bytes [0, n) are readable
//   win_begin, win_end : windows STARTING in [win_begin, win_end) are emitted
//                        (win_end is clamped to n-k+1 by the host)
// Internally positions are "aligned coordinates" a = p + shift where
// shift = data & 15, so every load is a 16-byte aligned block that contains at
// least one readable byte (never crosses into an unmapped page).
// ---------------------------------------------------------------------------
struct ScanGeom {
    const uint4* abase;  // data rounded down to 16 B
    uint64_t lo;         // first readable aligned coordinate  (= shift)
    uint64_t hi;         // one past last readable              (= shift + n)
    uint64_t wlo, whi;   // window-start range in aligned coordinates
    uint64_t g_begin, g_end;  // 512-byte groups that hold window starts
    int k;
};

__host__ __device__ inline ScanGeom kc_make_geom(const char* data, uint64_t n, uint64_t win_begin,
                                                 uint64_t win_end, int k) {
    ScanGeom g;
    const uint64_t addr = (uint64_t)(uintptr_t)data;
    const uint64_t shift = addr & 15u;
    g.abase = (const uint4*)(uintptr_t)(addr - shift);
    g.lo = shift;
    g.hi = shift + n;
    g.wlo = shift + win_begin;
    g.whi = shift + win_end;
    g.g_begin = g.wlo >> 9;
    g.g_end = (g.whi > g.wlo) ? ((g.whi + 511) >> 9) : g.g_begin;
    g.k = k;
    return g;
}

// load + decode block `b` (16-byte units) with out-of-buffer bytes marked bad
__device__ __forceinline__ Decoded16 kc_load_block(const ScanGeom& g, uint64_t b) {
    const uint64_t a0 = b << 4;
    Decoded16 d;
    if (a0 + 16 <= g.lo || a0 >= g.hi) {
        d.packed = 0;
        d.bad = 0xFFFFu;
        return d;
    }
    d = kc_decode16(kc_ldg_stream(g.abase + b));
    if (a0 < g.lo || a0 + 16 > g.hi) {
        const int l = a0 < g.lo ? (int)(g.lo - a0) : 0;
        const int h = a0 + 16 > g.hi ? (int)(g.hi - a0) : 16;
        const uint32_t keep = ((1u << h) - 1u) & ~((1u << l) - 1u);
        d.bad |= ~keep & 0xFFFFu;
    }
    return d;
}

// Split form for software pipelining: issue the 128-bit load early, decode when
// the data is needed (keeps two loads per lane in flight).
__device__ __forceinline__ uint4 kc_issue_block(const ScanGeom& g, uint64_t b) {
    const uint64_t a0 = b << 4;
    if (a0 + 16 <= g.lo || a0 >= g.hi) return make_uint4(0, 0, 0, 0);
    return kc_ldg_stream(g.abase + b);
}
__device__ __forceinline__ Decoded16 kc_finish_block(const ScanGeom& g, uint64_t b, uint4 raw) {
    const uint64_t a0 = b << 4;
    Decoded16 d = kc_decode16(raw);  // an all-zero block decodes to 16 bad bases
    if (a0 < g.lo || a0 + 16 > g.hi) {
        if (a0 + 16 <= g.lo || a0 >= g.hi) {
            d.bad = 0xFFFFu;
        } else {
            const int l = a0 < g.lo ? (int)(g.lo - a0) : 0;
            const int h = a0 + 16 > g.hi ? (int)(g.hi - a0) : 16;
            const uint32_t keep = ((1u << h) - 1u) & ~((1u << l) - 1u);
            d.bad |= ~keep & 0xFFFFu;
        }
    }
    return d;
}

// One lane's view of a 512-byte group: 16 own bases + up to 32 halo bases.
template <int HALO>
struct LaneWindow {
    uint32_t p0, p1, p2;  // packed bases: own, next 16, next-next 16 (HALO==2)
    uint32_t ok;          // bit j: window starting at own base j is valid and in range
    // 32-bit code (k <= 16) of the window starting at own base j
    __device__ __forceinline__ uint32_t code32(int j, uint32_t kmask) const {
        return __funnelshift_r(p0, p1, 2 * j) & kmask;
    }
    __device__ __forceinline__ uint64_t code64(int j, uint64_t kmask) const {
        const uint32_t lo = __funnelshift_r(p0, p1, 2 * j);
        const uint32_t hi = __funnelshift_r(p1, HALO == 2 ? p2 : 0u, 2 * j);
        return (((uint64_t)hi << 32) | lo) & kmask;
    }
};

// Drives a warp over groups [gb, ge) of the geometry; calls body(lw, a0) once
// per group per lane, a0 = aligned coordinate of the lane's first base.
// Pipeline: the block of group g+2 is requested at the top of the step for group
// g and decoded at its end, so one 128-bit load per lane is in flight during the
// whole body; groups g and g+1 are held decoded (g+1 feeds the halo of the last
// lanes).
template <int HALO, typename Body>
__device__ __forceinline__ void kc_warp_scan(const ScanGeom& g, uint64_t gb, uint64_t ge, Body body) {
    const int lane = threadIdx.x & 31;
    if (gb >= ge) return;
    Decoded16 cur = kc_load_block(g, gb * 32 + lane);
    Decoded16 nxt = kc_load_block(g, (gb + 1) * 32 + lane);
    for (uint64_t grp = gb; grp < ge; grp++) {
        const uint4 raw = kc_issue_block(g, (grp + 2) * 32 + lane);
        LaneWindow<HALO> lw;
        lw.p0 = cur.packed;
        uint32_t p1 = __shfl_down_sync(0xffffffffu, cur.packed, 1);
        uint32_t b1 = __shfl_down_sync(0xffffffffu, cur.bad, 1);
        const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
        const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
        if (lane == 31) {
            p1 = n0p;
            b1 = n0b;
        }
        lw.p1 = p1;
        uint64_t B = (uint64_t)cur.bad | ((uint64_t)b1 << 16);
        if (HALO == 2) {
            uint32_t p2 = __shfl_down_sync(0xffffffffu, cur.packed, 2);
            uint32_t b2 = __shfl_down_sync(0xffffffffu, cur.bad, 2);
            const uint32_t n1p = __shfl_sync(0xffffffffu, nxt.packed, 1);
            const uint32_t n1b = __shfl_sync(0xffffffffu, nxt.bad, 1);
            if (lane == 30) {
                p2 = n0p;
                b2 = n0b;
            } else if (lane == 31) {
                p2 = n1p;
                b2 = n1b;
            }
            lw.p2 = p2;
            B |= (uint64_t)b2 << 32;
        } else {
            lw.p2 = 0;
            B |= 0xFFFFull << 32;  // nothing beyond 32 bases is known
        }
        const uint64_t a0 = (grp * 32 + lane) << 4;
        uint32_t ok = ~(uint32_t)kc_window_bad(B, g.k) & 0xFFFFu;
        // window-start range (warp-uniform fast path for interior groups)
        if ((grp << 9) < g.wlo || ((grp + 1) << 9) > g.whi) {
            const int l = a0 < g.wlo ? (int)min((uint64_t)16, g.wlo - a0) : 0;
            const int h = a0 + 16 > g.whi ? (int)(g.whi > a0 ? g.whi - a0 : 0) : 16;
            ok &= ((1u << h) - 1u) & ~((1u << l) - 1u);
        }
        lw.ok = ok;
        body(lw, a0);
        // A body with data-dependent loops (staging, probing) leaves the lanes of a warp on separate paths, and nothing
        // brought them together before the next shuffle: ncu showed the decode below running with 5 of 32 threads per
        // instruction, 6.4 times as often as there are groups (sp_scatter_kernel, round 2).  Converge here.
        __syncwarp();
        cur = nxt;
        nxt = kc_finish_block(g, (grp + 2) * 32 + lane, raw);
    }
}
#endif  // __CUDACC__

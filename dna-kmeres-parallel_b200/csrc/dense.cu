// dense.cu — dense 4^k histogram counting (k <= 16), the hot path.
//
// Replaces the reference's count kernel sumKmereCoincidencesGlobalMemory
// (kernels.h:113-144: one CTA per sequence, one thread per 3-mer, each thread
// re-scanning the whole sequence byte by byte) with single-pass kernels:
//
//   dense_direct_kernel<SMEM>  k <= 7 : CTA-private 4^k uint32 bins in shared
//                                       memory, flushed once with global REDs.
//   dense_direct_kernel<GLOBAL> any k : one global RED per window (bins in L2).
//   dense_smem16_kernel         k = 8 : 65536 16-bit bins in shared memory with a
//                                       provably bounded spill to the global table.
//   partition path (k = 9..12, shown for 12): pass 1 groups A=5 consecutive windows
//       into one 32-bit "record" (the 16 bases they span), radix-partitions
//       records by 11 bits that all five windows share, staging them in shared
//       memory so every global write is a coalesced run; pass 2 gives each CTA
//       one partition and counts its five 8192-bin sub-tables with shared-memory
//       atomics, then adds them to the global table in 128-byte runs.
//       Random scatter to a 64 MiB table moves from L2 (<= 1 sector/clk/SM) into
//       shared memory (32 banks/clk/SM).  See DESIGN.md §3.3.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

// ---------------------------------------------------------------------------
// direct kernels
// ---------------------------------------------------------------------------
enum { BINS_SMEM = 0, BINS_GLOBAL = 1 };

template <int MODE>
__global__ void __launch_bounds__(256) dense_direct_kernel(ScanGeom g, uint32_t* __restrict__ table) {
    KC_DYN_SMEM(uint32_t, s_bins);
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
// k = 8: shared-memory-privatised bins with 16-bit counters
// ---------------------------------------------------------------------------
// 4^8 uint32 bins are 256 KB — more than one CTA can hold — so each CTA keeps the
// 65536 bins as 16-bit fields, two per 32-bit word (bin c lives in word c & 0x7FFF,
// half c >> 15): 128 KB.  A field must never wrap into its neighbour:
//   * the thread whose add makes a field reach a multiple of 0x4000 moves 0x4000 of
//     it to the global table (shared atomic add of -0x4000 on the field, one global
//     RED of +0x4000);
//   * the CTA meets at a barrier every 2 steps = 32768 adds.
// Invariant: every pending move was triggered at a value >= 0x4000 * (moves pending),
// so a field never goes negative; at a barrier nothing is pending and each field is
// < 0x4000 (every up-crossing of a multiple of 0x4000 is paired with one move down);
// between two barriers at most 32768 adds happen in the whole CTA, so a field stays
// below 0x4000 + 0x8000 = 0xC000 < 0x10000 whatever the scheduling.
// At the end every CTA writes its 32768 words to partials[cta][] and
// smem16_reduce_kernel adds them, unpacked, into the table.
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1)
dense_smem16_kernel(const uint4* __restrict__ base, uint64_t ngroups, uint32_t* __restrict__ table,
                    uint32_t* __restrict__ partials) {
    static_assert(DEPTH == 2, "the barrier period (2 steps = 32768 adds) is the unroll factor");
    KC_DYN_SMEM(uint32_t, words);  // 32768
    const uint32_t s_words = (uint32_t)__cvta_generic_to_shared(words);
    const int tid = threadIdx.x, lane = tid & 31;
    {
        uint4* w4 = reinterpret_cast<uint4*>(words);
        for (int i = tid; i < 32768 / 4; i += 1024) w4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint64_t nwarps = (uint64_t)gridDim.x * 32;
    const uint64_t units = ngroups / DEPTH;
    const uint64_t upw = (units + nwarps - 1) / nwarps;  // unrolled iterations of every warp (uniform)
    const uint64_t w = (uint64_t)blockIdx.x * 32 + (tid >> 5);
    const uint64_t gb = min(w * upw, units) * DEPTH;
    const uint32_t nsteps = (uint32_t)(min((w + 1) * upw, units) * DEPTH - gb);
    const uint4* ptr = base + gb * 32 + lane;
    uint4 raw[DEPTH];
    Decoded16 cur16;
    cur16.packed = 0;
    cur16.bad = 0xFFFFu;
    if (nsteps) {
        cur16 = kc_decode16(kc_ldg_stream(ptr));
#pragma unroll
        for (int q = 0; q < DEPTH; q++) raw[q] = kc_ldg_stream(ptr + 32 * (q + 1));
    }
    for (uint32_t i = 0; i < (uint32_t)upw * DEPTH; i += DEPTH) {
        if (i < nsteps) {  // warp-uniform; nsteps is a multiple of DEPTH
            uint4 fresh[DEPTH];
#pragma unroll
            for (int q = 0; q < DEPTH; q++) fresh[q] = kc_ldg_stream(ptr + 32 * (DEPTH + 1 + q));
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                const Decoded16 nxt = kc_decode16(raw[q]);
                uint32_t p1 = __shfl_down_sync(0xffffffffu, cur16.packed, 1);
                uint32_t b1 = __shfl_down_sync(0xffffffffu, cur16.bad, 1);
                const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
                const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
                if (lane == 31) {
                    p1 = n0p;
                    b1 = n0b;
                }
                const uint32_t p0 = cur16.packed;
                const uint32_t B32 = cur16.bad | (b1 << 16);
                uint32_t ok = 0xFFFFu;
                if (B32) ok = ~(uint32_t)kc_window_bad((uint64_t)B32 | (0xFFFFull << 32), 8) & 0xFFFFu;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (ok & (1u << j)) {
                        const uint32_t code = __funnelshift_r(p0, p1, 2 * j) & 0xFFFFu;
                        const uint32_t inc = (code >> 15) * 0xFFFFu + 1u;  // 1 or 0x10000
                        const uint32_t addr = s_words + (code & 0x7FFFu) * 4;
                        const uint32_t old = smem_atom_add(addr, inc);
                        if (((old + inc) & (inc * 0x3FFFu)) == 0) {  // the field reached 0x4000 or 0x8000
                            smem_atom_add(addr, 0u - inc * 0x4000u);
                            global_red_add(table + code, 0x4000u);
                        }
                    }
                }
                cur16 = nxt;
            }
            const uint32_t zero = kc_opaque_zero(cur16.bad);  // bad < 2^16: pins the copies here
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                raw[q].x = fresh[q].x | zero;
                raw[q].y = fresh[q].y | zero;
                raw[q].z = fresh[q].z | zero;
                raw[q].w = fresh[q].w | zero;
            }
            ptr += 32 * DEPTH;
        }
        __syncthreads();  // at most 32768 adds per CTA between two barriers
    }
    {
        const uint4* w4 = reinterpret_cast<const uint4*>(words);
        uint4* out = reinterpret_cast<uint4*>(partials + (uint64_t)blockIdx.x * 32768);
        for (int i = tid; i < 32768 / 4; i += 1024) out[i] = w4[i];
    }
}

__global__ void smem16_reduce_kernel(const uint32_t* __restrict__ partials, int nparts, uint32_t* __restrict__ table) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32768) return;
    uint32_t lo = 0, hi = 0;
    for (int p = 0; p < nparts; p++) {
        const uint32_t v = partials[(uint64_t)p * 32768 + w];
        lo += v & 0xFFFFu;
        hi += v >> 16;
    }
    table[w] += lo;
    table[w + 32768] += hi;
}

// ---------------------------------------------------------------------------
// k = 8, CHECKSUM variant (KC_DENSE_SMEM16C; first B200 run pending, not the default).
// The kernel above pays for its exactness inside the hot loop: a RETURNING shared atomic,
// a compare and a branch per window, plus a CTA barrier every two steps (it runs at 4.2
// windows/clk/SM where shared memory sustains 9).  Here the hot loop is decode + one
// non-returning `red.shared.add` per window and nothing else; a 16-bit field that wraps
// is found AFTERWARDS:
//   every add puts +1 into one 16-bit field, so without a wrap  sum(fields) == adds of the CTA;
//   a low field that wraps loses 65536 and gives +1 to the high field (sum -65535), a high
//   field that wraps loses 65536 (sum -65536); c1 and c2 such events change the sum by
//   -(65535 c1 + 65536 c2), which is 0 only for c1 = c2 = 0.
// A CTA whose checksum fails raises flags[cta]; the reduce kernel skips its partial table
// and smem16_repair_kernel recounts exactly that CTA's groups with global REDs.  Uniform
// data never fails (3.1 Gbp: ~320 hits per bin and CTA); poly-A fails everywhere and is
// still exact, at the direct path's speed.
// ---------------------------------------------------------------------------
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1)
dense_smem16c_kernel(const uint4* __restrict__ base, uint64_t ngroups, uint32_t* __restrict__ partials,
                     uint32_t* __restrict__ flags) {
    KC_DYN_SMEM(uint32_t, words);  // 32768
    __shared__ unsigned long long s_red[64];
    const uint32_t s_words = (uint32_t)__cvta_generic_to_shared(words);
    const int tid = threadIdx.x, lane = tid & 31;
    {
        uint4* w4 = reinterpret_cast<uint4*>(words);
        for (int i = tid; i < 32768 / 4; i += 1024) w4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint64_t nwarps = (uint64_t)gridDim.x * 32;
    const uint64_t units = ngroups / DEPTH;
    const uint64_t upw = (units + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * 32 + (tid >> 5);
    const uint64_t gb = min(w * upw, units) * DEPTH;
    const uint32_t nsteps = (uint32_t)(min((w + 1) * upw, units) * DEPTH - gb);
    const uint4* ptr = base + gb * 32 + lane;
    uint32_t nadds = 0;
    if (nsteps) {
        uint4 raw[DEPTH];
        Decoded16 cur16 = kc_decode16(kc_ldg_stream(ptr));
#pragma unroll
        for (int q = 0; q < DEPTH; q++) raw[q] = kc_ldg_stream(ptr + 32 * (q + 1));
        for (uint32_t i = 0; i < nsteps; i += DEPTH) {
            uint4 fresh[DEPTH];
#pragma unroll
            for (int q = 0; q < DEPTH; q++) fresh[q] = kc_ldg_stream(ptr + 32 * (DEPTH + 1 + q));
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                const Decoded16 nxt = kc_decode16(raw[q]);
                uint32_t p1 = __shfl_down_sync(0xffffffffu, cur16.packed, 1);
                uint32_t b1 = __shfl_down_sync(0xffffffffu, cur16.bad, 1);
                const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
                const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
                if (lane == 31) {
                    p1 = n0p;
                    b1 = n0b;
                }
                const uint32_t p0 = cur16.packed;
                const uint32_t B32 = cur16.bad | (b1 << 16);
                uint32_t ok = 0xFFFFu;
                if (B32) ok = ~(uint32_t)kc_window_bad((uint64_t)B32 | (0xFFFFull << 32), 8) & 0xFFFFu;
                nadds += (uint32_t)__popc(ok);
                if (ok == 0xFFFFu) {  // the common case carries no per-window predicate
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        // code << 2 straight from the funnel shifter: word address = bits [2,17), half = bit 17
                        const uint32_t c4 = j ? __funnelshift_r(p0, p1, 2 * j - 2) : (p0 << 2);
                        smem_red_add(s_words + (c4 & 0x1FFFCu), (c4 & 0x20000u) ? 0x10000u : 1u);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if (ok & (1u << j)) {
                            const uint32_t c4 = j ? __funnelshift_r(p0, p1, 2 * j - 2) : (p0 << 2);
                            smem_red_add(s_words + (c4 & 0x1FFFCu), (c4 & 0x20000u) ? 0x10000u : 1u);
                        }
                    }
                }
                cur16 = nxt;
            }
            const uint32_t zero = kc_opaque_zero(cur16.bad);  // bad < 2^16: pins the copies here
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                raw[q].x = fresh[q].x | zero;
                raw[q].y = fresh[q].y | zero;
                raw[q].z = fresh[q].z | zero;
                raw[q].w = fresh[q].w | zero;
            }
            ptr += 32 * DEPTH;
        }
    }
    __syncthreads();
    // checksum: sum of all 16-bit fields vs the adds this CTA made
    // (per-warp sums fit 32 bits: a CTA makes < 2^32 adds; REDUX does each in one instruction)
    uint32_t f32 = 0;
    for (int i = tid; i < 32768; i += 1024) {
        const uint32_t v = words[i];
        f32 += (v & 0xFFFFu) + (v >> 16);
    }
    const unsigned long long fsum = __reduce_add_sync(0xffffffffu, f32), asum = __reduce_add_sync(0xffffffffu, nadds);
    if (lane == 0) {
        s_red[tid >> 5] = fsum;
        s_red[32 + (tid >> 5)] = asum;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long f = 0, a = 0;
        for (int q = 0; q < 32; q++) {
            f += s_red[q];
            a += s_red[32 + q];
        }
        flags[blockIdx.x] = (f == a) ? 0u : 1u;
        s_red[0] = (f == a) ? 0ull : 1ull;
    }
    __syncthreads();
    if (s_red[0] == 0) {
        const uint4* w4 = reinterpret_cast<const uint4*>(words);
        uint4* out = reinterpret_cast<uint4*>(partials + (uint64_t)blockIdx.x * 32768);
        for (int i = tid; i < 32768 / 4; i += 1024) out[i] = w4[i];
    }
}

// grid (32768 / 256, SLICES): slice y adds the partial tables p = y, y + SLICES, ... (148 sequential loads per thread
// made this kernel 27 us for 19 MB in round 2's first capture; 8 slices and one RED per word and slice instead)
__global__ void smem16c_reduce_kernel(const uint32_t* __restrict__ partials, const uint32_t* __restrict__ flags,
                                      int nparts, uint32_t* __restrict__ table) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= 32768) return;
    uint32_t lo = 0, hi = 0;
    for (int p = blockIdx.y; p < nparts; p += gridDim.y) {
        if (flags[p]) continue;  // that CTA's table wrapped: smem16_repair_kernel recounts its groups
        const uint32_t v = kc_ld_cg(partials + (uint64_t)p * 32768 + w);
        lo += v & 0xFFFFu;
        hi += v >> 16;
    }
    if (lo) atomicAdd(table + w, lo);
    if (hi) atomicAdd(table + w + 32768, hi);
}

// one CTA per pass-1 CTA; does nothing unless that CTA's checksum failed
template <int DEPTH>
__global__ void __launch_bounds__(256)
smem16_repair_kernel(const uint4* __restrict__ base, uint64_t ngroups, int nparts, const uint32_t* __restrict__ flags,
                     uint32_t* __restrict__ table) {
    if (flags[blockIdx.x] == 0) return;
    if (threadIdx.x == 0) KC_STAT(6);
    const uint64_t nwarps = (uint64_t)nparts * 32;
    const uint64_t units = ngroups / DEPTH;
    const uint64_t upw = (units + nwarps - 1) / nwarps;
    const uint64_t gb = min((uint64_t)blockIdx.x * 32 * upw, units) * DEPTH;        // the failed CTA's 32 warps
    const uint64_t ge = min(((uint64_t)blockIdx.x + 1) * 32 * upw, units) * DEPTH;  // own one contiguous run
    ScanGeom g;
    g.abase = base;
    g.lo = 0;
    g.hi = (ngroups + 1) << 9;  // interior groups: the group behind the last one is readable too (host)
    g.wlo = gb << 9;
    g.whi = ge << 9;
    g.g_begin = gb;
    g.g_end = ge;
    g.k = 8;
    const uint64_t ng = ge - gb;
    const uint64_t per = (ng + 7) >> 3;
    const int warp = threadIdx.x >> 5;
    kc_warp_scan<1>(g, gb + min((uint64_t)warp * per, ng), gb + min((uint64_t)(warp + 1) * per, ng),
                    [&](const LaneWindow<1>& lw, uint64_t) {
                        const uint32_t ok = lw.ok & 0xFFFFu;
                        if (ok == 0) return;
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if (ok & (1u << j)) atomicAdd(&table[lw.code32(j, 0xFFFFu)], 1u);
                    });
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
__device__ __noinline__ void part_fallback(uint32_t rec, uint32_t okbits, uint32_t* table) {
#pragma unroll
    for (int r = 0; r < C::A; r++)
        if (okbits & (1u << r)) global_red_add(table + part_code<C>(rec, r), 1u);
}

// Pass 1.  Each warp owns a contiguous run of 512-byte groups and never waits for
// another warp: no CTA barrier and no global atomic in the main loop.  The kernel
// only sees INTERIOR groups: every block it touches (including DEPTH+1 groups of
// prefetch past its last group) is readable and every window it can emit lies in
// the requested range, so the hot loop has no bounds checks, no range masks and a
// single code path; the host sends the few windows before/after the interior to
// dense_direct_kernel.
//
// Loads: every iteration of the DEPTH-times unrolled loop first issues the DEPTH
// 128-bit loads the NEXT iteration will decode and copies them into the ring at its
// end, so DEPTH loads per lane are in flight for DEPTH steps — the bare scan runs
// at the HBM peak this way (tools/microbench2.cu).  (ptxas turns any "load into the
// ring slot in place" formulation into load-to-spare-register + MOV right behind
// it, and the MOV waits for the load: measured 3x slower.)
//
// Shared memory per CTA: state[P] (low 16 bits = slots reserved, high 16 bits =
// slots written), cur[P] (absolute word offset of the next chunk in this CTA's
// private global region of partition p), buf[P][CAP] (staged records; CAP = 24 =
// one 96-byte chunk of three full 32-byte sectors).
//
// Staging protocol (all shared-memory, CTA scope):
//   writer : slot = atom.add(state[p], 1) & 0xFFFF;
//            slot <  CAP : buf[p][slot] = rec; w = atom.add(state[p], 1<<16) >> 16;
//                          the writer that makes w == CAP-1 has seen every slot
//                          written: that LANE flushes the bin by itself;
//            slot >= CAP : the bin is full and its flush is in flight (a window of
//                          a few dozen cycles) -> count the record's windows with
//                          global REDs instead (well under 1 % of records).
//   flusher: one lane reads buf[p][0..CAP) with 128-bit shared loads, resets
//            state[p] = 0 (one store clears both halves, so a new generation can
//            start immediately), advances cur[p] and writes the chunk with 128-bit
//            global stores.  Several lanes of a warp flush different bins in the
//            same divergent region, so the cost per step does not grow with their
//            number (the warp-cooperative flush of an earlier version did).
// Shared-memory requests of a thread are performed in program order, so a writer's
// store precedes its "written" increment and the flusher's loads precede its
// reset; any record a later generation stores therefore lands after the loads.
// Every record is consumed exactly once: by a chunk, by the final flush, or by
// the RED fallback.  Parity against the oracle at 3.1 G windows checks this.
//
// Global layout: slabs[cta][p][region_cap] — a CTA writes into one contiguous
// area, the (cta, p) regions need no global cursor; counts[p][cta] is published
// at the end.
template <typename C, int THREADS, int MINB, int DEPTH, int ABLATE = 0>
__global__ void __launch_bounds__(THREADS, MINB)
part_scatter_kernel(const uint4* __restrict__ base, uint64_t ngroups, uint32_t* __restrict__ table,
                    uint32_t* __restrict__ slabs, uint32_t* __restrict__ counts, uint32_t region_cap) {
    static_assert(C::CAP % 8 == 0, "chunks are moved 128 bits at a time, in two halves");
    constexpr int NW = THREADS / 32;
    KC_DYN_SMEM(uint32_t, smem);
    const uint32_t s_state = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_cur = s_state + C::P * 4;
    const uint32_t s_buf = s_state + 2 * C::P * 4;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int b = tid; b < C::P; b += THREADS) {
        smem[b] = 0;
        smem[C::P + b] = b * region_cap;
    }
    __syncthreads();

    // ngroups is a multiple of DEPTH (host) and so is every warp's share: the
    // unrolled loop below then needs no per-step guard.  (A guard makes the ring
    // registers conditionally assigned; ptxas then loads into a spare register and
    // inserts a MOV at the join that waits for the load — no overlap at all.)
    const uint64_t nwarps = (uint64_t)gridDim.x * NW;
    const uint64_t units = ngroups / DEPTH;
    const uint64_t upw = (units + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * NW + (tid >> 5);
    const uint64_t gb = min(w * upw, units) * DEPTH;
    const uint32_t nsteps = (uint32_t)(min((w + 1) * upw, units) * DEPTH - gb);
    uint32_t* const my_slabs = slabs + (uint64_t)blockIdx.x * C::P * region_cap;  // < 2^32 words per CTA

    // flush bin b (this lane completed it).  All shared loads are issued first and the
    // bin is reopened right behind them, before the global side starts: the window in
    // which other warps find the bin full (and fall back to REDs) is a handful of issue
    // slots instead of an atomic round trip plus a store burst.
    auto flush_bin = [&](uint32_t b) {
        constexpr int Q = C::CAP / 4;  // 128-bit pieces
        const uint32_t src = s_buf + b * (C::CAP * 4);
        // (Two flush variants were measured on B200 and removed: one TMA bulk copy per bin, 2.77 ms against 2.63 —
        // the bin stays closed until the unit has read it; 256-bit stores, no gain.  BENCH_r01 config.probe.)
        uint4 v[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) v[q] = smem_ld128(src + 16 * q);
        smem_st(s_state + b * 4, 0u);  // every slot has been read: the bin is free again
        const uint32_t pos = smem_atom_add(s_cur + b * 4, (uint32_t)C::CAP);
        if (pos + C::CAP <= (b + 1) * region_cap) {
            uint4* dst = reinterpret_cast<uint4*>(my_slabs + pos);  // pos, region_cap: multiples of 4 words
#pragma unroll
            for (int q = 0; q < Q; q++) dst[q] = v[q];
        } else {  // region full (skewed input): rare, slow, exact (static indices: v stays in registers)
#pragma unroll
            for (int q = 0; q < Q; q++) {
                part_fallback<C>(v[q].x, C::AMASK, table);
                part_fallback<C>(v[q].y, C::AMASK, table);
                part_fallback<C>(v[q].z, C::AMASK, table);
                part_fallback<C>(v[q].w, C::AMASK, table);
            }
        }
    };

    uint32_t pend = 0;       // DEFER variant: a record waiting for its second attempt
    bool have_pend = false;
    if (nsteps) {
        // record starts are the positions = 0 mod A counted from the first interior byte
        int d = (int)(((gb % C::A) * (512 % C::A) + (uint32_t)lane * (16 % C::A)) % C::A);
        const uint4* ptr = base + gb * 32 + lane;
        uint4 raw[DEPTH];
        Decoded16 cur16 = kc_decode16(kc_ldg_stream(ptr));
#pragma unroll
        for (int q = 0; q < DEPTH; q++) raw[q] = kc_ldg_stream(ptr + 32 * (q + 1));

        for (uint32_t i = 0; i < nsteps; i += DEPTH) {
            // issue the loads of the NEXT iteration's groups first (i+DEPTH+1 ...): they
            // have this whole iteration (DEPTH steps) to land before the copy at its end
            uint4 fresh[DEPTH];
#pragma unroll
            for (int q = 0; q < DEPTH; q++) fresh[q] = kc_ldg_stream(ptr + 32 * (DEPTH + 1 + q));
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                {
                    const Decoded16 nxt = kc_decode16(raw[q]);  // group i+q+1, loaded one iteration ago
                    uint32_t p1 = __shfl_down_sync(0xffffffffu, cur16.packed, 1);
                    uint32_t b1 = __shfl_down_sync(0xffffffffu, cur16.bad, 1);
                    const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
                    const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
                    if (lane == 31) {
                        p1 = n0p;
                        b1 = n0b;
                    }
                    const uint32_t p0 = cur16.packed;
                    const uint32_t B32 = cur16.bad | (b1 << 16);
                    uint32_t ok = (1u << (33 - C::K)) - 1u;  // the windows that fit the 32 known bases
                    if (B32) ok = ~(uint32_t)kc_window_bad((uint64_t)B32 | (0xFFFFull << 32), C::K);
                    const int j0 = d ? C::A - d : 0;  // first record start among this lane's 16 positions
                    // records at j = j0, j0+A, ... while j < 16 (sh = 2j < 32)
                    constexpr int NSLOT = (16 + C::A - 1) / C::A;
#pragma unroll
                    for (int t = 0; t < NSLOT; t++) {
                        const uint32_t sh = 2 * (uint32_t)j0 + 2 * C::A * t;
                        const uint32_t okr = (ok >> (j0 + C::A * t)) & C::AMASK;
                        if (sh < 32 && okr) {
                            const uint32_t rec = __funnelshift_r(p0, p1, sh);
                            if (okr != C::AMASK) {
                                part_fallback<C>(rec, okr, table);  // rare: next to an N run
                            } else if (ABLATE == 3) {
                                // DEFER variant (not the default until measured): a record that meets a
                                // full bin waits in a register for this lane's next record slot, where it
                                // gets ONE more attempt before it is counted with REDs.  Motive
                                // (profiles/r01_ncu_per_source_line.txt): 2.2 % of the records take the
                                // RED fallback, but that divergent 42-instruction function runs in 30 % of
                                // the warp iterations and costs 21 % of the kernel's instructions; a bin
                                // reopens within a few hundred cycles, a lane's next slot comes later.
                                uint32_t r = have_pend ? pend : rec;
                                bool retry = have_pend;
                                have_pend = false;
                                for (;;) {
                                    const uint32_t pid = (r >> C::KB0) & (C::P - 1);
                                    const uint32_t slot = smem_atom_add(s_state + pid * 4, 1u) & 0xFFFFu;
                                    if (slot < (uint32_t)C::CAP) {
                                        smem_st(s_buf + (pid * C::CAP + slot) * 4, r);
                                        const uint32_t wr = smem_atom_add(s_state + pid * 4, 0x10000u) >> 16;
                                        if (wr == (uint32_t)C::CAP - 1) flush_bin(pid);
                                    } else if (retry) {
                                        KC_STAT(2);
                                        part_fallback<C>(r, C::AMASK, table);
                                    } else {
                                        KC_STAT(1);
                                        pend = r;
                                        have_pend = true;
                                    }
                                    if (!retry) break;
                                    retry = false;
                                    r = rec;
                                }
                            } else {
                                const uint32_t pid = (rec >> C::KB0) & (C::P - 1);
                                const uint32_t slot = smem_atom_add(s_state + pid * 4, 1u) & 0xFFFFu;
                                if (slot < (uint32_t)C::CAP) {
                                    smem_st(s_buf + (pid * C::CAP + slot) * 4, rec);
                                    const uint32_t wr = smem_atom_add(s_state + pid * 4, 0x10000u) >> 16;
                                    if (wr == (uint32_t)C::CAP - 1) flush_bin(pid);
                                } else {
                                    KC_STAT(0);
                                    part_fallback<C>(rec, C::AMASK, table);
                                }
                            }
                        }
                    }
                    cur16 = nxt;
                    d += 512 % C::A;
                    if (d >= C::A) d -= C::A;
                }
            }
            // Copy into the ring at the END of the iteration.  ptxas would hoist plain
            // copies to just behind the last read of raw[q] (and stall there on the
            // load); OR-ing in a zero it cannot see through, produced after the last
            // shared-memory operation of the iteration, pins them here.
            const uint32_t zero = kc_opaque_zero((uint32_t)d);  // d in [0, A)
#pragma unroll
            for (int q = 0; q < DEPTH; q++) {
                raw[q].x = fresh[q].x | zero;
                raw[q].y = fresh[q].y | zero;
                raw[q].z = fresh[q].z | zero;
                raw[q].w = fresh[q].w | zero;
            }
            ptr += 32 * DEPTH;
        }
    }
    if (ABLATE == 3 && have_pend) part_fallback<C>(pend, C::AMASK, table);
    // final flush of the partially filled bins, then publish the region lengths
    __syncthreads();
    for (int b = tid >> 5; b < C::P; b += NW) {
        const uint32_t c = smem[b] & 0xFFFFu;  // < CAP: a full bin was flushed by its last writer
        const uint32_t base_off = b * region_cap;
        const uint32_t pos = smem[C::P + b] - base_off;
        uint32_t stored = pos < region_cap ? pos : region_cap;  // chunks are a prefix
        if ((uint32_t)lane < c) {
            const uint32_t r = smem[2 * C::P + b * C::CAP + lane];
            if (pos + C::CAP <= region_cap)
                my_slabs[base_off + pos + lane] = r;
            else
                part_fallback<C>(r, C::AMASK, table);
        }
        if (pos + C::CAP <= region_cap) stored = pos + c;
        if (lane == 0) counts[(uint64_t)b * gridDim.x + blockIdx.x] = stored;
    }
}

// Pass 2.  One partition at a time per CTA (dynamic queue); NBINS uint32 bins in
// shared memory; the partition's records lie in `nregions` private regions (one
// per pass-1 CTA) which the 32 warps take round-robin.
template <typename C>
__global__ void __launch_bounds__(1024, (C::NBINS * 4 <= 100 * 1024) ? 2 : 1)
part_count_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slabs,
                  const uint32_t* __restrict__ counts, uint32_t region_cap, uint32_t nregions,
                  uint32_t* __restrict__ work_counter) {
    KC_DYN_SMEM(uint32_t, bins);
    __shared__ uint32_t s_part;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
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
        auto count_rec = [&](uint32_t rec) {
            // Y = record with the key bits removed; window r's bin = Y bits [2r, 2r+KB0)
            const uint32_t Y = (rec & SUBMASK) | ((rec >> (2 * C::K)) << C::KB0);
#pragma unroll
            for (int r = 0; r < C::A; r++) atomicAdd(&bins[r * C::SUB + ((Y >> (2 * r)) & SUBMASK)], 1u);
        };
        for (uint32_t reg = warp; reg < nregions; reg += 32) {
            const uint32_t n = counts[(uint64_t)part * nregions + reg];
            const uint32_t* src = slabs + ((uint64_t)reg * C::P + part) * region_cap;  // chunk aligned
            const uint4* src4 = reinterpret_cast<const uint4*>(src);
            const uint32_t n4 = n >> 2;
            uint32_t i = lane;
            for (; i + 32 < n4; i += 64) {  // two loads in flight per lane
                const uint4 v0 = kc_ldg_stream(src4 + i);
                const uint4 v1 = kc_ldg_stream(src4 + i + 32);
                count_rec(v0.x);
                count_rec(v0.y);
                count_rec(v0.z);
                count_rec(v0.w);
                count_rec(v1.x);
                count_rec(v1.y);
                count_rec(v1.z);
                count_rec(v1.w);
            }
            for (; i < n4; i += 32) {
                const uint4 v = kc_ldg_stream(src4 + i);
                count_rec(v.x);
                count_rec(v.y);
                count_rec(v.z);
                count_rec(v.w);
            }
            for (uint32_t t = (n4 << 2) + lane; t < n; t += 32) count_rec(src[t]);
        }
        __syncthreads();
        // add the sub-tables to the global table: for alignment r the bin is
        //   low (KB0-2r bits) | key << (KB0-2r) | high (2r bits) << (2K-2r)
#pragma unroll
        for (int r = 0; r < C::A; r++) {
            const int lowbits = C::KB0 - 2 * r;
            const uint32_t keypart = part << lowbits;
            for (int f = tid; f < C::SUB; f += 1024) {
                const uint32_t v = bins[r * C::SUB + f];
                if (v) {
                    const uint32_t low = (uint32_t)f & ((1u << lowbits) - 1u);
                    const uint32_t high = (uint32_t)f >> lowbits;
                    global_red_add(table + (low | keypart | (high << (2 * C::K - 2 * r))), v);
                }
            }
        }
        __syncthreads();
    }
}

using Part12 = PartCfg<12, 5, 11, 16>;

// Pass 2, PAIRED variant for k = 12 (KC_DENSE_PARTITION_PAIR; first B200 run pending, not the
// default).  The count kernel above is bound by shared-memory wavefronts: one random atomic per
// window, 3.35 wavefronts per 32-lane atomic (profiles/r01_ncu_instruction_mix.txt).  A record
// holds 5 consecutive windows; windows (0,1) are the two 12-mers of the 13-mer at offset 0 and
// (2,3) those of the 13-mer at offset 2, so counting the two 13-mers and window 4 is THREE
// increments per record instead of five, and the 13-mer counts are folded into 12-mer bins while
// the sub-tables are flushed (the number of global REDs stays 40960 per partition):
//   T0[32768]  13-mer at offset 0, index = 13-mer code minus the 11 key bits   16-bit fields
//   T2[32768]  13-mer at offset 2                                               16-bit fields
//   T4[8192]   window 4 (as in the kernel above)                                32-bit
// = 160 KB, the same shared memory.  16-bit fields can wrap on skewed input; as in
// dense_smem16c_kernel a wrap is found afterwards (sum of all fields != 2 x records of the
// partition) and the CTA then recounts the partition with the five 32-bit tables.
__global__ void __launch_bounds__(1024, 1)
part_count_pair12_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slabs,
                         const uint32_t* __restrict__ counts, uint32_t region_cap, uint32_t nregions,
                         uint32_t* __restrict__ work_counter) {
    using C = Part12;
    KC_DYN_SMEM(uint32_t, bins);  // 40960 words
    __shared__ uint32_t s_part, s_bad;
    __shared__ unsigned long long s_red[64];
    const uint32_t s_bins = (uint32_t)__cvta_generic_to_shared(bins);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t W0 = 0, W2 = 16384, W4 = 32768;  // word offsets of T0, T2, T4
    for (;;) {
        if (tid == 0) s_part = atomicAdd(work_counter, 1u);
        {
            uint4* b4 = reinterpret_cast<uint4*>(bins);
            for (int i = tid; i < 40960 / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= (uint32_t)C::P) break;
        uint32_t nrec = 0;
        auto count_rec = [&](uint32_t rec) {
            const uint32_t c0 = rec & 0x3FFFFFFu;          // bases 0..12
            const uint32_t c2 = (rec >> 4) & 0x3FFFFFFu;   // bases 2..14
            const uint32_t c4 = rec >> 8;                  // bases 4..15
            const uint32_t i0 = (c0 & 0x1FFFu) | ((c0 >> 24) << 13);  // key = record bits [13,24) removed: 15 bits
            const uint32_t i2 = (c2 & 0x1FFu) | ((c2 >> 20) << 9);    // 15 bits
            const uint32_t i4 = (c4 & 0x1Fu) | ((c4 >> 16) << 5);     // 13 bits
            smem_red_add(s_bins + (W0 + (i0 & 0x3FFFu)) * 4, (i0 & 0x4000u) ? 0x10000u : 1u);
            smem_red_add(s_bins + (W2 + (i2 & 0x3FFFu)) * 4, (i2 & 0x4000u) ? 0x10000u : 1u);
            smem_red_add(s_bins + (W4 + i4) * 4, 1u);
            nrec++;
        };
        auto count_rec_classic = [&](uint32_t rec) {  // the five 32-bit sub-tables of part_count_kernel
            const uint32_t Y = (rec & 0x1FFFu) | ((rec >> 24) << 13);
#pragma unroll
            for (int r = 0; r < 5; r++) smem_red_add(s_bins + (r * 8192 + ((Y >> (2 * r)) & 0x1FFFu)) * 4, 1u);
        };
        auto for_each_record = [&](auto&& f) {
            for (uint32_t reg = warp; reg < nregions; reg += 32) {
                const uint32_t n = counts[(uint64_t)part * nregions + reg];
                const uint32_t* src = slabs + ((uint64_t)reg * C::P + part) * region_cap;  // chunk aligned
                const uint4* src4 = reinterpret_cast<const uint4*>(src);
                const uint32_t n4 = n >> 2;
                uint32_t i = lane;
                for (; i + 32 < n4; i += 64) {  // two loads in flight per lane
                    const uint4 v0 = kc_ldg_stream(src4 + i);
                    const uint4 v1 = kc_ldg_stream(src4 + i + 32);
                    f(v0.x);
                    f(v0.y);
                    f(v0.z);
                    f(v0.w);
                    f(v1.x);
                    f(v1.y);
                    f(v1.z);
                    f(v1.w);
                }
                for (; i < n4; i += 32) {
                    const uint4 v = kc_ldg_stream(src4 + i);
                    f(v.x);
                    f(v.y);
                    f(v.z);
                    f(v.w);
                }
                for (uint32_t t = (n4 << 2) + lane; t < n; t += 32) f(src[t]);
            }
        };
        for_each_record(count_rec);
        __syncthreads();
        // checksum: the 16-bit fields of T0 and T2 must add up to two increments per record
        uint32_t f32 = 0;  // per-warp sums fit 32 bits (a partition holds < 2^31 records); one REDUX each
        for (int i = tid; i < 32768; i += 1024) {
            const uint32_t v = bins[i];
            f32 += (v & 0xFFFFu) + (v >> 16);
        }
        const unsigned long long fsum = __reduce_add_sync(0xffffffffu, f32), asum = __reduce_add_sync(0xffffffffu, 2u * nrec);
        if (lane == 0) {
            s_red[warp] = fsum;
            s_red[32 + warp] = asum;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long f = 0, a = 0;
            for (int q = 0; q < 32; q++) {
                f += s_red[q];
                a += s_red[32 + q];
            }
            s_bad = (f == a) ? 0u : 1u;
#ifdef KC_EMU  // a real wrap needs > 65535 records of one partition in one bin (0.7 GB of input): the
               // emulator tests force the recount path instead (the detection itself is the one
               // dense_smem16c_kernel's tests exercise with real wraps)
            if (getenv("KC_EMU_FORCE_PAIR_RECOUNT") && (part & 1u)) s_bad = 1u;
#endif
        }
        __syncthreads();
        if (s_bad) {  // a field wrapped (skewed input): recount this partition with 32-bit bins
            KC_STAT(7);
            {
                uint4* b4 = reinterpret_cast<uint4*>(bins);
                for (int i = tid; i < 40960 / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
            for_each_record(count_rec_classic);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const int lowbits = 13 - 2 * r;
                for (int f = tid; f < 8192; f += 1024) {
                    const uint32_t v = bins[r * 8192 + f];
                    if (v) {
                        const uint32_t low = (uint32_t)f & ((1u << lowbits) - 1u), high = (uint32_t)f >> lowbits;
                        global_red_add(table + (low | (part << lowbits) | (high << (24 - 2 * r))), v);
                    }
                }
            }
        } else {
            // fold and flush.  A field f of a 16-bit table lives in word f & 0x3FFF, half f >> 14.
            // T0, f = low13 | h << 13 (h = the 13-mer's last base):
            //   window 0 = low13 | key << 13                      (sum over h)
            //   window 1 = low13 >> 2 | key << 11 | h << 22       (sum over the 13-mer's first base)
            // T2, f = low9 | hi6 << 9:
            //   window 2 = low9 | key << 9 | (hi6 & 15) << 20     (sum over hi6 >> 4)
            //   window 3 = low9 >> 2 | key << 7 | hi6 << 18       (sum over low9 & 3)
            for (int x = tid; x < 8192; x += 1024) {
                const uint32_t a = bins[W0 + x], b = bins[W0 + 8192 + x];
                const uint32_t v0 = (a & 0xFFFFu) + (a >> 16) + (b & 0xFFFFu) + (b >> 16);
                if (v0) global_red_add(table + ((uint32_t)x | (part << 13)), v0);
                const uint32_t c = bins[W2 + x], d = bins[W2 + 8192 + x];
                const uint32_t v2 = (c & 0xFFFFu) + (c >> 16) + (d & 0xFFFFu) + (d >> 16);
                if (v2) global_red_add(table + (((uint32_t)x & 0x1FFu) | (part << 9) | (((uint32_t)x >> 9) << 20)), v2);
                const uint32_t v4 = bins[W4 + x];
                if (v4) global_red_add(table + (((uint32_t)x & 0x1Fu) | (part << 5) | (((uint32_t)x >> 5) << 16)), v4);
            }
            for (int y = tid; y < 4096; y += 1024) {  // y = word block (4 words) of a 16-bit table
                const uint4 w = *reinterpret_cast<const uint4*>(bins + W0 + 4 * y);
                const uint32_t lo = (w.x & 0xFFFFu) + (w.y & 0xFFFFu) + (w.z & 0xFFFFu) + (w.w & 0xFFFFu);
                const uint32_t hi = (w.x >> 16) + (w.y >> 16) + (w.z >> 16) + (w.w >> 16);
                // words 4y..4y+3 hold fields f = 4y + t (half 0) and f + 16384 (half 1); f = low13 | h << 13
                const uint32_t q = (uint32_t)y & 2047u, hb = (uint32_t)y >> 11;  // low13 >> 2, h & 1
                if (lo) global_red_add(table + (q | (part << 11) | (hb << 22)), lo);
                if (hi) global_red_add(table + (q | (part << 11) | ((hb + 2u) << 22)), hi);
                const uint4 u = *reinterpret_cast<const uint4*>(bins + W2 + 4 * y);
                const uint32_t lo2 = (u.x & 0xFFFFu) + (u.y & 0xFFFFu) + (u.z & 0xFFFFu) + (u.w & 0xFFFFu);
                const uint32_t hi2 = (u.x >> 16) + (u.y >> 16) + (u.z >> 16) + (u.w >> 16);
                // fields f = 4y + t = low9 | hi6 << 9 with hi6 < 32 (half 0) and hi6 + 32 (half 1)
                const uint32_t q2 = (uint32_t)y & 127u, h6 = (uint32_t)y >> 7;  // low9 >> 2, hi6 & 31
                if (lo2) global_red_add(table + (q2 | (part << 7) | (h6 << 18)), lo2);
                if (hi2) global_red_add(table + (q2 | (part << 7) | ((h6 + 32u) << 18)), hi2);
            }
        }
        __syncthreads();
    }
}

// Pass 2, TWO increments per record (KC_DENSE_PARTITION_TRIO; first B200 run pending): the same
// idea one step further.  The 14-mer at offset 0 contains windows 0, 1, 2 and the 13-mer at offset 3
// windows 3, 4:
//   T14[131072]  14-mer at offset 0 minus the 11 key bits, 8-bit fields   (128 KB)
//   T13[32768]   13-mer at offset 3 minus the key bits,   16-bit fields   ( 64 KB)
// A partition of the 3.1 Gbp genome puts 300 K records into 131072 fields (2.3 per field), so 8 bits
// are plenty for unskewed data; a wrapped field changes the sum of all fields by -255 / -256 (8-bit)
// or -65535 / -65536 (16-bit), never by 0, so  sum(T14) == records && sum(T13) == records  proves
// that nothing wrapped; otherwise the partition is recounted with the five 32-bit tables.
// Fold at the flush (f = field index, key = partition):
//   T14, f = low13 | hi4 << 13 (hi4 = bases 12,13); word f & 0x7FFF, byte f >> 15
//     window 0 = low13 | key << 13                         sum over hi4            (16 fields)
//     window 1 = low13 >> 2 | key << 11 | (hi4 & 3) << 22  sum over low13 & 3, hi4 >> 2
//     window 2 = low13 >> 4 | key << 9 | hi4 << 20         sum over low13 & 15
//   T13, f = low7 | hi8 << 7; word f & 0x3FFF, half f >> 14
//     window 3 = low7 | key << 7 | (hi8 & 63) << 18        sum over hi8 >> 6
//     window 4 = low7 >> 2 | key << 5 | hi8 << 16          sum over low7 & 3
__global__ void __launch_bounds__(1024, 1)
part_count_trio12_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slabs,
                         const uint32_t* __restrict__ counts, uint32_t region_cap, uint32_t nregions,
                         uint32_t* __restrict__ work_counter) {
    using C = Part12;
    KC_DYN_SMEM(uint32_t, bins);  // 49152 words: T14 = words [0, 32768), T13 = words [32768, 49152)
    __shared__ uint32_t s_part, s_bad;
    __shared__ unsigned long long s_red[96];
    const uint32_t s_bins = (uint32_t)__cvta_generic_to_shared(bins);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t W13 = 32768, NWORDS = 49152;
    auto bytesum = [](uint32_t v) { return (v & 0xFFu) + ((v >> 8) & 0xFFu) + ((v >> 16) & 0xFFu) + (v >> 24); };
    auto bytesum4 = [](uint4 w, int sh) {  // the byte at bit `sh` of four words
        return ((w.x >> sh) & 0xFFu) + ((w.y >> sh) & 0xFFu) + ((w.z >> sh) & 0xFFu) + ((w.w >> sh) & 0xFFu);
    };
    for (;;) {
        if (tid == 0) s_part = atomicAdd(work_counter, 1u);
        {
            uint4* b4 = reinterpret_cast<uint4*>(bins);
            for (int i = tid; i < (int)NWORDS / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= (uint32_t)C::P) break;
        uint32_t nrec = 0;
        auto count_rec = [&](uint32_t rec) {
            const uint32_t c14 = rec & 0xFFFFFFFu;  // bases 0..13
            const uint32_t c13 = rec >> 6;          // bases 3..15
            const uint32_t i0 = (c14 & 0x1FFFu) | ((c14 >> 24) << 13);  // key = record bits [13,24) removed: 17 bits
            const uint32_t i3 = (c13 & 0x7Fu) | ((c13 >> 18) << 7);     // 15 bits
            smem_red_add(s_bins + (i0 & 0x7FFFu) * 4, 1u << (8 * (i0 >> 15)));
            smem_red_add(s_bins + (W13 + (i3 & 0x3FFFu)) * 4, (i3 & 0x4000u) ? 0x10000u : 1u);
            nrec++;
        };
        auto count_rec_classic = [&](uint32_t rec) {  // the five 32-bit sub-tables of part_count_kernel
            const uint32_t Y = (rec & 0x1FFFu) | ((rec >> 24) << 13);
#pragma unroll
            for (int r = 0; r < 5; r++) smem_red_add(s_bins + (r * 8192 + ((Y >> (2 * r)) & 0x1FFFu)) * 4, 1u);
        };
        auto for_each_record = [&](auto&& f) {
            for (uint32_t reg = warp; reg < nregions; reg += 32) {
                const uint32_t n = counts[(uint64_t)part * nregions + reg];
                const uint32_t* src = slabs + ((uint64_t)reg * C::P + part) * region_cap;  // chunk aligned
                const uint4* src4 = reinterpret_cast<const uint4*>(src);
                const uint32_t n4 = n >> 2;
                uint32_t i = lane;
                for (; i + 32 < n4; i += 64) {  // two loads in flight per lane
                    const uint4 v0 = kc_ldg_stream(src4 + i);
                    const uint4 v1 = kc_ldg_stream(src4 + i + 32);
                    f(v0.x);
                    f(v0.y);
                    f(v0.z);
                    f(v0.w);
                    f(v1.x);
                    f(v1.y);
                    f(v1.z);
                    f(v1.w);
                }
                for (; i < n4; i += 32) {
                    const uint4 v = kc_ldg_stream(src4 + i);
                    f(v.x);
                    f(v.y);
                    f(v.z);
                    f(v.w);
                }
                for (uint32_t t = (n4 << 2) + lane; t < n; t += 32) f(src[t]);
            }
        };
        for_each_record(count_rec);
        __syncthreads();
        // checksums: every record put one increment into each table
        uint32_t a14 = 0, a13 = 0;  // per-warp sums fit 32 bits; one REDUX each
        for (int i = tid; i < (int)W13; i += 1024) a14 += bytesum(bins[i]);
        for (int i = tid; i < 16384; i += 1024) {
            const uint32_t v = bins[W13 + i];
            a13 += (v & 0xFFFFu) + (v >> 16);
        }
        const unsigned long long s14 = __reduce_add_sync(0xffffffffu, a14), s13 = __reduce_add_sync(0xffffffffu, a13),
                                 asum = __reduce_add_sync(0xffffffffu, nrec);
        if (lane == 0) {
            s_red[warp] = s14;
            s_red[32 + warp] = s13;
            s_red[64 + warp] = asum;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long a = 0, b = 0, c = 0;
            for (int q = 0; q < 32; q++) {
                a += s_red[q];
                b += s_red[32 + q];
                c += s_red[64 + q];
            }
            s_bad = (a == c && b == c) ? 0u : 1u;
#ifdef KC_EMU  // see part_count_pair12_kernel
            if (getenv("KC_EMU_FORCE_PAIR_RECOUNT") && (part & 1u)) s_bad = 1u;
#endif
        }
        __syncthreads();
        if (s_bad) {  // a field wrapped (skewed input): recount this partition with 32-bit bins
            KC_STAT(7);
            {
                uint4* b4 = reinterpret_cast<uint4*>(bins);
                for (int i = tid; i < 40960 / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
            for_each_record(count_rec_classic);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const int lowbits = 13 - 2 * r;
                for (int f = tid; f < 8192; f += 1024) {
                    const uint32_t v = bins[r * 8192 + f];
                    if (v) {
                        const uint32_t low = (uint32_t)f & ((1u << lowbits) - 1u), high = (uint32_t)f >> lowbits;
                        global_red_add(table + (low | (part << lowbits) | (high << (24 - 2 * r))), v);
                    }
                }
            }
        } else {
            for (int x = tid; x < 8192; x += 1024) {
                // window 0: all 16 fields low13 = x (4 words x + 8192 b12, 4 bytes each)
                const uint32_t v0 = bytesum(bins[x]) + bytesum(bins[x + 8192]) + bytesum(bins[x + 16384]) + bytesum(bins[x + 24576]);
                if (v0) global_red_add(table + ((uint32_t)x | (part << 13)), v0);
                // window 3: low7 | (hi8 & 63) << 7 = x; the four hi8 >> 6 are the two halves of words x and x + 8192
                const uint32_t c = bins[W13 + x], d = bins[W13 + 8192 + x];
                const uint32_t v3 = (c & 0xFFFFu) + (c >> 16) + (d & 0xFFFFu) + (d >> 16);
                if (v3) global_red_add(table + (((uint32_t)x & 0x7Fu) | (part << 7) | (((uint32_t)x >> 7) << 18)), v3);
            }
            for (int y = tid; y < 8192; y += 1024) {
                // window 1: word block y of T14 = fields 4q + t (t < 4) with b12 = y >> 11; all four bytes (b13)
                const uint4 w = *reinterpret_cast<const uint4*>(bins + 4 * y);
                const uint32_t v1 = bytesum(w.x) + bytesum(w.y) + bytesum(w.z) + bytesum(w.w);
                if (v1) global_red_add(table + (((uint32_t)y & 2047u) | (part << 11) | (((uint32_t)y >> 11) << 22)), v1);
            }
            for (int z = tid; z < 2048; z += 1024) {
                // window 2: 16 consecutive words (low13 & 15) of block z = q4 | b12 << 9; one bin per byte (b13)
                const uint4* p4 = reinterpret_cast<const uint4*>(bins + 16 * z);
                const uint4 a = p4[0], b = p4[1], c = p4[2], d = p4[3];
                const uint32_t q4 = (uint32_t)z & 511u, b12 = (uint32_t)z >> 9;
#pragma unroll
                for (int b13 = 0; b13 < 4; b13++) {
                    const uint32_t v2 = bytesum4(a, 8 * b13) + bytesum4(b, 8 * b13) + bytesum4(c, 8 * b13) + bytesum4(d, 8 * b13);
                    if (v2) global_red_add(table + (q4 | (part << 9) | ((b12 | ((uint32_t)b13 << 2)) << 20)), v2);
                }
            }
            for (int y = tid; y < 4096; y += 1024) {
                // window 4: word block y of T13 = fields low7 = 4q + t, hi8 & 127 = y >> 5; halves = hi8 >> 7
                const uint4 u = *reinterpret_cast<const uint4*>(bins + W13 + 4 * y);
                const uint32_t lo = (u.x & 0xFFFFu) + (u.y & 0xFFFFu) + (u.z & 0xFFFFu) + (u.w & 0xFFFFu);
                const uint32_t hi = (u.x >> 16) + (u.y >> 16) + (u.z >> 16) + (u.w >> 16);
                const uint32_t q = (uint32_t)y & 31u, h7 = (uint32_t)y >> 5;
                if (lo) global_red_add(table + (q | (part << 5) | (h7 << 16)), lo);
                if (hi) global_red_add(table + (q | (part << 5) | ((h7 + 128u) << 16)), hi);
            }
        }
        __syncthreads();
    }
}

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
        KC_LAUNCH(dense_direct_kernel<BINS_SMEM>, grid, 256, smem, st, g, d_table);
    } else {
        KC_LAUNCH(dense_direct_kernel<BINS_GLOBAL>, grid, 256, 0, st, g, d_table);
    }
    KC_LAUNCH_CHECK(ctx, "dense_direct_kernel");
    if (ctx->timing) {
        KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
        ctx->timed_kernels = 1;
    }
    return KC_OK;
}

// launch shape of pass 1
template <typename C, int THREADS_, int MINB_, int DEPTH_>
struct ScatterShape {
    using Cfg = C;
    static constexpr int THREADS = THREADS_, MINB = MINB_, DEPTH = DEPTH_;
};

static int dense_direct(kc_ctx* ctx, const ScanGeom& g, uint32_t* d_table, cudaStream_t st);

template <typename S>
static int dense_partition(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                           uint32_t* d_table, cudaStream_t st, bool defer, int pair = 0) {
    using C = typename S::Cfg;
    const ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, C::K);
    // interior groups [G0, G1): fully readable, all their windows requested, and
    // DEPTH+1 groups of prefetch past G1 still readable
    const uint64_t G0 = (max(g.lo, g.wlo) + 511) >> 9;
    const uint64_t lim = min(g.hi, g.whi);
    const uint64_t Gl = lim >> 9;
    const uint64_t G1 = Gl > (uint64_t)(2 * S::DEPTH + 2) ? Gl - (2 * S::DEPTH + 2) : 0;
    if (G1 <= G0 + 64) return dense_direct(ctx, g, d_table, st);
    const uint64_t ngroups = (G1 - G0) / S::DEPTH * S::DEPTH;  // every warp runs whole unrolled iterations
    const uint64_t nrec = (ngroups * 512 + C::A - 1) / C::A;  // records start at (G0<<9) + A*m
    const uint64_t shift = g.lo;                              // aligned coordinate of byte 0
    const uint64_t head_end = (G0 << 9) - shift;              // first window the interior kernel owns
    const uint64_t tail_begin = (G0 << 9) + nrec * C::A - shift;

    constexpr int NW = S::THREADS / 32;
    const uint64_t want = (ngroups + NW - 1) / NW;
    const uint64_t maxg = (uint64_t)ctx->sm_count * S::MINB;
    const int grid1 = (int)(want > maxg ? maxg : want);
    // private region of every (partition, pass-1 CTA): mean + 12.5 % + slack, a whole
    // number of 128-byte lines.  Overflow falls back to global REDs.
    uint64_t cap = nrec / ((uint64_t)C::P * grid1);
    cap = cap + cap / 8 + 4 * C::CAP;
    constexpr uint64_t kUnit = (C::CAP % 32 == 0) ? C::CAP : (C::CAP % 16 == 0 ? 2 * C::CAP : 4 * C::CAP);
    cap = (cap + kUnit - 1) / kUnit * kUnit;  // whole chunks and whole 128-byte lines
    if ((uint64_t)C::P * cap >= (1ull << 32)) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition region too large");
    const size_t nregions = (size_t)C::P * grid1;
    const size_t ctl_bytes = (nregions + 64) * sizeof(uint32_t);  // counts[P][grid1] + work counter
    const size_t ctl_pad = (ctl_bytes + 255) & ~(size_t)255;
    const size_t slab_bytes = nregions * cap * sizeof(uint32_t);
    int rc = kc_scratch_reserve(ctx, ctl_pad + slab_bytes);
    if (rc) return rc;
    uint32_t* counts = (uint32_t*)ctx->scratch;
    uint32_t* work = counts + nregions;
    uint32_t* slabs = (uint32_t*)((char*)ctx->scratch + ctl_pad);
    KC_CUDA(ctx, cudaMemsetAsync(work, 0, 64 * sizeof(uint32_t), st));

    const size_t smem1 = (size_t)(2 * C::P + C::P * C::CAP) * sizeof(uint32_t);
    const size_t smem2 = (size_t)C::NBINS * sizeof(uint32_t);
    const uint4* base = g.abase + (G0 << 5);
    static const int ablate_env = getenv("KC_PART_ABLATE") ? atoi(getenv("KC_PART_ABLATE")) : 0;  // 3 = deferred retry (tests)
    const int ablate = (defer || ablate_env == 3) ? 3 : 0;
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
#define KC_LAUNCH_SCATTER(ABL)                                                                              \
    do {                                                                                                    \
        auto kern = part_scatter_kernel<C, S::THREADS, S::MINB, S::DEPTH, ABL>;                   \
        KC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));  \
        KC_LAUNCH(kern, grid1, S::THREADS, smem1, st, base, ngroups, d_table, slabs, counts, (uint32_t)cap);       \
    } while (0)
    if (ablate == 3)
        KC_LAUNCH_SCATTER(3);
    else
        KC_LAUNCH_SCATTER(0);
#undef KC_LAUNCH_SCATTER
    KC_LAUNCH_CHECK(ctx, "part_scatter_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
    const int ctas2 = (smem2 + 1024) * 2 <= ctx->smem_optin + 1024 && smem2 <= 100 * 1024 ? 2 : 1;
    int grid2 = ctx->sm_count * ctas2 < C::P ? ctx->sm_count * ctas2 : C::P;
    static const int pair_env = getenv("KC_PART_PAIR") ? atoi(getenv("KC_PART_PAIR")) : 0;  // measurement aid
    const int pair_mode = pair ? pair : pair_env;  // 1 = pairs (3 increments per record), 2 = 14-mer + 13-mer (2 increments)
    if (pair_mode == 2 && std::is_same<C, Part12>::value) {
        const size_t smem3 = 49152 * sizeof(uint32_t);
        KC_CUDA(ctx, cudaFuncSetAttribute(part_count_trio12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        KC_LAUNCH(part_count_trio12_kernel, grid2, 1024, smem3, st, d_table, slabs, counts, (uint32_t)cap, (uint32_t)grid1, work);
        KC_LAUNCH_CHECK(ctx, "part_count_trio12_kernel");
    } else if (pair_mode && std::is_same<C, Part12>::value) {
        KC_CUDA(ctx, cudaFuncSetAttribute(part_count_pair12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        KC_LAUNCH(part_count_pair12_kernel, grid2, 1024, smem2, st, d_table, slabs, counts, (uint32_t)cap, (uint32_t)grid1, work);
        KC_LAUNCH_CHECK(ctx, "part_count_pair12_kernel");
    } else {
        KC_CUDA(ctx, cudaFuncSetAttribute(part_count_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        KC_LAUNCH(part_count_kernel<C>, grid2, 1024, smem2, st, d_table, slabs, counts, (uint32_t)cap, (uint32_t)grid1, work);
        KC_LAUNCH_CHECK(ctx, "part_count_kernel");
    }
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[2], st));
    // the windows before and after the interior
    const bool timing = ctx->timing;
    ctx->timing = false;
    rc = KC_OK;
    if (head_end > win_begin) rc = dense_direct(ctx, kc_make_geom(d_data, nbytes, win_begin, head_end, C::K), d_table, st);
    if (!rc && tail_begin < win_end) rc = dense_direct(ctx, kc_make_geom(d_data, nbytes, tail_begin, win_end, C::K), d_table, st);
    ctx->timing = timing;
    if (timing) ctx->timed_kernels = 2;
    return rc;
}

// k = 8 path: interior groups through dense_smem16_kernel (or its checksum variant), the rest
// through dense_direct
static int dense_smem16(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                        uint32_t* d_table, cudaStream_t st, bool checksum) {
    constexpr int DEPTH = 2;
    const ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, 8);
    const uint64_t G0 = (max(g.lo, g.wlo) + 511) >> 9;
    const uint64_t lim = min(g.hi, g.whi);
    const uint64_t Gl = lim >> 9;
    const uint64_t G1 = Gl > (uint64_t)(2 * DEPTH + 2) ? Gl - (2 * DEPTH + 2) : 0;
    if (G1 <= G0 + 64) return dense_direct(ctx, g, d_table, st);
    const uint64_t ngroups = (G1 - G0) / DEPTH * DEPTH;
    const uint64_t shift = g.lo;
    const uint64_t head_end = (G0 << 9) - shift;
    const uint64_t tail_begin = ((G0 + ngroups) << 9) - shift;  // every window start of the interior groups is counted
    const uint64_t want = (ngroups / DEPTH + 31) / 32;
    const int grid = (int)(want > (uint64_t)ctx->sm_count ? (uint64_t)ctx->sm_count : want);
    int rc = kc_scratch_reserve(ctx, ((size_t)grid * 32768 + 1024) * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t* partials = (uint32_t*)ctx->scratch;
    uint32_t* flags = partials + (size_t)grid * 32768;  // checksum variant: one word per CTA (grid <= 1024)
    const size_t smem = 32768 * sizeof(uint32_t);
    const uint4* base = g.abase + (G0 << 5);
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
    if (checksum) {
        KC_CUDA(ctx, cudaFuncSetAttribute(dense_smem16c_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KC_LAUNCH(dense_smem16c_kernel<DEPTH>, grid, 1024, smem, st, base, ngroups, partials, flags);
        KC_LAUNCH_CHECK(ctx, "dense_smem16c_kernel");
        if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
        KC_LAUNCH(smem16c_reduce_kernel, dim3(32768 / 256, 8), 256, 0, st, partials, flags, grid, d_table);
        KC_LAUNCH_CHECK(ctx, "smem16c_reduce_kernel");
        KC_LAUNCH(smem16_repair_kernel<DEPTH>, grid, 256, 0, st, base, ngroups, grid, flags, d_table);
        KC_LAUNCH_CHECK(ctx, "smem16_repair_kernel");
    } else {
        KC_CUDA(ctx, cudaFuncSetAttribute(dense_smem16_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KC_LAUNCH(dense_smem16_kernel<DEPTH>, grid, 1024, smem, st, base, ngroups, d_table, partials);
        KC_LAUNCH_CHECK(ctx, "dense_smem16_kernel");
        if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
        KC_LAUNCH(smem16_reduce_kernel, 32768 / 256, 256, 0, st, partials, grid, d_table);
        KC_LAUNCH_CHECK(ctx, "smem16_reduce_kernel");
    }
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[2], st));
    const bool timing = ctx->timing;
    ctx->timing = false;
    rc = KC_OK;
    if (head_end > win_begin) rc = dense_direct(ctx, kc_make_geom(d_data, nbytes, win_begin, head_end, 8), d_table, st);
    if (!rc && tail_begin < win_end) rc = dense_direct(ctx, kc_make_geom(d_data, nbytes, tail_begin, win_end, 8), d_table, st);
    ctx->timing = timing;
    if (timing) ctx->timed_kernels = 2;
    return rc;
}

int kc_dense_direct_range(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end, int k,
                          uint32_t* d_table, cudaStream_t st) {
    return dense_direct(ctx, kc_make_geom(d_data, nbytes, win_begin, win_end, k), d_table, st);
}


static uint64_t g_partition_min_windows = 1ull << 26;  // below this the direct path wins

extern "C" int kc_count_dense_range_async(kc_ctx* ctx, const char* d_data, uint64_t nbytes,
                                          uint64_t win_begin, uint64_t win_end, int k,
                                          uint32_t* d_table, int algo, void* stream) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!d_table || (!d_data && nbytes)) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    if (algo != KC_DENSE_AUTO && algo != KC_DENSE_DIRECT && algo != KC_DENSE_PARTITION && algo != KC_DENSE_SMEM16C &&
        algo != KC_DENSE_PARTITION_DEFER && algo != KC_DENSE_PARTITION_PAIR && algo != KC_DENSE_PARTITION_TRIO &&
        algo != KC_DENSE_PARTITION_WIDE && algo != KC_DENSE_PARTITION_DEFER_PAIR && algo != KC_DENSE_PARTITION_DEFER_TRIO &&
        algo != KC_DENSE_PARTITION_WIDE2)
        return kc_set_error(ctx, KC_ERR_INVALID, "unknown dense algo %d", algo);
    if (algo == KC_DENSE_SMEM16C) {
        if (k != 8) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "KC_DENSE_SMEM16C is the k = 8 path (k=%d)", k);
        DeviceGuard dg8(ctx->device);
        if (nbytes < 8) return KC_OK;
        if (win_end > nbytes - 7) win_end = nbytes - 7;
        if (win_begin >= win_end) return KC_OK;
        return dense_smem16(ctx, d_data, nbytes, win_begin, win_end, d_table, (cudaStream_t)stream, true);
    }
    if (algo == KC_DENSE_PARTITION_WIDE) {
        if (k != 12) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "KC_DENSE_PARTITION_WIDE is built for k = 12 (k=%d)", k);
        DeviceGuard dgw(ctx->device);
        if (nbytes < 12) return KC_OK;
        if (win_end > nbytes - 11) win_end = nbytes - 11;
        if (win_begin >= win_end) return KC_OK;
        return kc_dense_partition_wide(ctx, d_data, nbytes, win_begin, win_end, d_table, (cudaStream_t)stream);
    }
    // KC_DENSE_AUTO at k = 12: the second-generation scatter with seven windows per record, by measurement on B200
    // (3.1 Gbp: 2.75 ms against 3.48 ms for the five-window scatter + two-increment count; profiles/r02_*)
    static const bool auto_old12 = getenv("KC_DENSE_AUTO_R01") != nullptr;  // measurement aid: round 1's choice
    if (algo == KC_DENSE_AUTO && k == 12 && !auto_old12 && nbytes >= 12) {
        const uint64_t we = win_end > nbytes - 11 ? nbytes - 11 : win_end;
        if (we > win_begin && we - win_begin >= g_partition_min_windows) algo = KC_DENSE_PARTITION_WIDE2;
    }
    if (algo == KC_DENSE_PARTITION_WIDE2) {
        if (k != 12) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "KC_DENSE_PARTITION_WIDE2 is built for k = 12 (k=%d)", k);
        DeviceGuard dgw(ctx->device);
        if (nbytes < 12) return KC_OK;
        if (win_end > nbytes - 11) win_end = nbytes - 11;
        if (win_begin >= win_end) return KC_OK;
        return kc_dense_partition_wide2(ctx, d_data, nbytes, win_begin, win_end, d_table, (cudaStream_t)stream);
    }
    const bool defer = (algo == KC_DENSE_PARTITION_DEFER || algo == KC_DENSE_PARTITION_DEFER_PAIR || algo == KC_DENSE_PARTITION_DEFER_TRIO);
    // KC_DENSE_AUTO at k = 12 takes the two-increment count (measured on B200, 3.1 Gbp: count pass 1.33 -> 0.83 ms,
    // BENCH_r01 config.probe); KC_DENSE_PARTITION stays the five-sub-table count so that the two can be compared.
    const int pair = (algo == KC_DENSE_PARTITION_PAIR || algo == KC_DENSE_PARTITION_DEFER_PAIR) ? 1
                     : (algo == KC_DENSE_PARTITION_TRIO || algo == KC_DENSE_PARTITION_DEFER_TRIO || (algo == KC_DENSE_AUTO && k == 12)) ? 2
                                                                                                                                        : 0;
    if (pair && k != 12) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "KC_DENSE_PARTITION_PAIR/TRIO are built for k = 12 (k=%d)", k);
    if (defer || (pair && algo != KC_DENSE_AUTO)) algo = KC_DENSE_PARTITION;
    DeviceGuard dg(ctx->device);
    if (nbytes < (uint64_t)k) return KC_OK;
    const uint64_t nwin = nbytes - k + 1;
    if (win_end > nwin) win_end = nwin;
    if (win_begin >= win_end) return KC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, k);
    const bool can_part = (k >= 9 && k <= 12);
    if (algo == KC_DENSE_PARTITION && !can_part)
        return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition path is built for k = 9..12 (k=%d)", k);
    const bool use_part =
        can_part && (algo == KC_DENSE_PARTITION ||
                     (algo == KC_DENSE_AUTO && (win_end - win_begin) >= g_partition_min_windows));
    // k = 8: the checksum variant is the default (B200, config 2: 0.103 -> 0.074 ms, profiles/r02_*)
    if (k == 8 && algo == KC_DENSE_AUTO && (win_end - win_begin) >= (1ull << 22))
        return dense_smem16(ctx, d_data, nbytes, win_begin, win_end, d_table, st, true);
    if (use_part) {
        // PartCfg<K, A, KB, CAP>: A windows per 32-bit record (K + A - 1 <= 16 bases), KB key bits
        // inside the bases all A windows share, CAP records per chunk.  k = 12 shape measured on
        // B200 (profiles/r01_scatter_shapes.txt): CAP 16 / 3 loads in flight is the fastest.
        static const int shape = getenv("KC_PART_SHAPE") ? atoi(getenv("KC_PART_SHAPE")) : 0;  // tuning aid
        switch (k) {
            case 12:
                if (shape == 1) return dense_partition<ScatterShape<PartCfg<12, 5, 11, 24>, 1024, 1, 2>>(ctx, d_data, nbytes, win_begin, win_end, d_table, st, defer);
                return dense_partition<ScatterShape<Part12, 1024, 1, 3>>(ctx, d_data, nbytes, win_begin, win_end, d_table, st, defer, pair);
            case 11: return dense_partition<ScatterShape<PartCfg<11, 6, 11, 16>, 1024, 1, 3>>(ctx, d_data, nbytes, win_begin, win_end, d_table, st, defer);
            case 10: return dense_partition<ScatterShape<PartCfg<10, 7, 8, 16>, 1024, 1, 3>>(ctx, d_data, nbytes, win_begin, win_end, d_table, st, defer);
            default: return dense_partition<ScatterShape<PartCfg<9, 6, 8, 16>, 1024, 1, 3>>(ctx, d_data, nbytes, win_begin, win_end, d_table, st, defer);
        }
    }
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

// End to end from host memory: chunked H2D on the copy stream, counting on the compute stream, events for the
// hand-off.  Chunk c is counted for the windows that END inside it (so only bytes already on the device are touched);
// the device buffer holds the whole input.  h_table != NULL: the table is copied to the host; the device table stays in
// ctx->scratch2 (its first 4 * 4^k bytes) either way, which is what the hybrid host path (packed.cu) adds up.
int kc_dense_host_plain(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* h_table) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if (!h_data && nbytes) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
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
    ctx->last_h2d_bytes = nbytes;
    if (h_table) KC_CUDA(ctx, cudaMemcpyAsync(h_table, d_table, table_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

extern "C" int kc_count_dense_host(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* h_table) {
    if (ctx && !h_table) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    return kc_dense_host_plain(ctx, h_data, nbytes, k, h_table);
}

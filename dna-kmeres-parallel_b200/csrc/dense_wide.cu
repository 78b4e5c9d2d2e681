// dense_wide.cu — KC_DENSE_PARTITION_WIDE: the k = 12 partition path with SEVEN windows per record
// (first B200 run pending; written without a GPU, verified on the CPU emulator, never picked by
// KC_DENSE_AUTO until measured).
//
// Why.  Pass 1 of the shipped path (dense.cu) is 2/3 of the step and its cost is per RECORD: two
// shared atomics, one store, 1/16 of a bin flush.  All windows of a record must share the 11-bit
// partition key; k-mers that start A-1 positions apart share 12-A+1 bases, so A <= 7.  The shipped
// path stops at A = 5 only because 5 windows span 16 bases = the 32 bits of a slab record.  But the
// key bits of a record are the same for every record of a partition: they need not be stored.
// Seven windows span 18 bases = 36 bits; minus the 11 key bits = 25 bits, still one 32-bit slab
// record.  Records per base drop from 1/5 to 1/7 (-29 % staging work, slab traffic and flushes).
//
// Pass 2 cannot afford seven 32 KB sub-tables, and does not need them (cf. part_count_trio12_kernel):
//   T0  the 14-mer at offset 0 (windows 0,1,2)   131072 fields of 4 bits   64 KB
//   T3  the 14-mer at offset 3 (windows 3,4,5)   131072 fields of 4 bits   64 KB
//   T6  window 6                                   8192 fields of 16 bits  16 KB
// Three increments per seven windows.  With the key bits removed, Y = the slab record, and the three
// indices are plain shifts of Y:  Y & 0x1FFFF,  (Y >> 6) & 0x1FFFF,  Y >> 12.  A partition of the
// 3.1 Gbp genome holds 216 K records (1.65 per 4-bit field); a field that wraps (15 -> 0 carries
// into its neighbour) changes the sum of all fields by -15 or -16, so  sum(T0) == sum(T3) ==
// sum(T6) == records  proves that nothing wrapped; otherwise the partition is recounted with seven
// 32-bit sub-tables (224 KB, the largest thing a CTA can hold).
//
// Record layout (bases of the record, 2 bits each, little-endian like every code here):
//   bits [0,13) low | [13,24) KEY | [24,36) high        Y = low | high << 13   (25 bits)
#include "common.cuh"

namespace {

constexpr int WK = 12, WA = 7, WP = 2048, WCAP = 16;

__device__ __forceinline__ uint64_t wide_rec(uint32_t Y, uint32_t part) {  // 36-bit record from slab record + key
    return (uint64_t)(Y & 0x1FFFu) | ((uint64_t)part << 13) | ((uint64_t)(Y >> 13) << 24);
}

// count the valid windows of a record directly (records next to an N run, records that met a
// full bin twice, region overflow)
__device__ __noinline__ void wide_fallback(uint64_t rec, uint32_t okbits, uint32_t* table) {
#pragma unroll
    for (int r = 0; r < WA; r++)
        if (okbits & (1u << r)) global_red_add(table + (uint32_t)((rec >> (2 * r)) & 0xFFFFFFu), 1u);
}

// Pass 1: as part_scatter_kernel (interior groups only, loads issued one unrolled iteration ahead,
// barrier-free staging, per-lane 128-bit flush) with 18-base records, the deferred retry of
// KC_PART_ABLATE=3 built in, and slab records that omit the key.
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1)
part_scatter7_kernel(const uint4* __restrict__ base, uint64_t ngroups, uint32_t* __restrict__ table,
                     uint32_t* __restrict__ slabs, uint32_t* __restrict__ counts, uint32_t region_cap) {
    KC_DYN_SMEM(uint32_t, smem);
    const uint32_t s_state = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t s_cur = s_state + WP * 4;
    const uint32_t s_buf = s_state + 2 * WP * 4;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int b = tid; b < WP; b += 1024) {
        smem[b] = 0;
        smem[WP + b] = b * region_cap;
    }
    __syncthreads();
    const uint64_t nwarps = (uint64_t)gridDim.x * 32;
    const uint64_t units = ngroups / DEPTH;
    const uint64_t upw = (units + nwarps - 1) / nwarps;
    const uint64_t w = (uint64_t)blockIdx.x * 32 + (tid >> 5);
    const uint64_t gb = min(w * upw, units) * DEPTH;
    const uint32_t nsteps = (uint32_t)(min((w + 1) * upw, units) * DEPTH - gb);
    uint32_t* const my_slabs = slabs + (uint64_t)blockIdx.x * WP * region_cap;

    auto flush_bin = [&](uint32_t b) {
        const uint32_t src = s_buf + b * (WCAP * 4);
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = smem_ld128(src + 16 * q);
        smem_st(s_state + b * 4, 0u);
        const uint32_t pos = smem_atom_add(s_cur + b * 4, (uint32_t)WCAP);
        if (pos + WCAP <= (b + 1) * region_cap) {
            uint4* dst = reinterpret_cast<uint4*>(my_slabs + pos);
#pragma unroll
            for (int q = 0; q < 4; q++) dst[q] = v[q];
        } else {  // region full (skewed input): rare, slow, exact
#pragma unroll
            for (int q = 0; q < 4; q++) {
                wide_fallback(wide_rec(v[q].x, b), 0x7Fu, table);
                wide_fallback(wide_rec(v[q].y, b), 0x7Fu, table);
                wide_fallback(wide_rec(v[q].z, b), 0x7Fu, table);
                wide_fallback(wide_rec(v[q].w, b), 0x7Fu, table);
            }
        }
    };

    uint32_t pendY = 0, pendP = 0;  // a record waiting for its second attempt (deferred retry)
    bool have_pend = false;
    if (nsteps) {
        int d = (int)(((gb % WA) * (512 % WA) + (uint32_t)lane * (16 % WA)) % WA);
        const uint4* ptr = base + gb * 32 + lane;
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
                // the 32 bases behind this lane's 16: lanes +1 and +2, or lanes 0/1 of the next group
                uint32_t p1 = __shfl_down_sync(0xffffffffu, cur16.packed, 1);
                uint32_t b1 = __shfl_down_sync(0xffffffffu, cur16.bad, 1);
                uint32_t p2 = __shfl_down_sync(0xffffffffu, cur16.packed, 2);
                uint32_t b2 = __shfl_down_sync(0xffffffffu, cur16.bad, 2);
                const uint32_t n0p = __shfl_sync(0xffffffffu, nxt.packed, 0);
                const uint32_t n0b = __shfl_sync(0xffffffffu, nxt.bad, 0);
                const uint32_t n1p = __shfl_sync(0xffffffffu, nxt.packed, 1);
                const uint32_t n1b = __shfl_sync(0xffffffffu, nxt.bad, 1);
                if (lane == 31) {
                    p1 = n0p;
                    b1 = n0b;
                    p2 = n1p;
                    b2 = n1b;
                } else if (lane == 30) {
                    p2 = n0p;
                    b2 = n0b;
                }
                const uint32_t p0 = cur16.packed;
                // windows starting at this lane's 16 bases + 6 more: they end within the 48 known bases
                uint32_t ok = 0x3FFFFFu;
                if (cur16.bad | b1 | b2) {
                    const uint64_t B = (uint64_t)cur16.bad | ((uint64_t)b1 << 16) | ((uint64_t)b2 << 32) | (0xFFFFull << 48);
                    ok = ~(uint32_t)kc_window_bad(B, WK) & 0x3FFFFFu;
                }
                const int j0 = d ? WA - d : 0;
                constexpr int NSLOT = (16 + WA - 1) / WA;
#pragma unroll
                for (int t = 0; t < NSLOT; t++) {
                    const int j = j0 + WA * t;
                    const uint32_t okr = (ok >> j) & 0x7Fu;
                    if (j < 16 && okr) {
                        const uint32_t lo = __funnelshift_r(p0, p1, 2 * j);          // bases j .. j+15
                        const uint32_t hi = __funnelshift_r(p1, p2, 2 * j) & 0xFu;   // bases j+16, j+17
                        if (okr != 0x7Fu) {
                            wide_fallback((uint64_t)lo | ((uint64_t)hi << 32), okr, table);  // rare: next to an N run
                        } else {
                            const uint32_t recY = (lo & 0x1FFFu) | ((lo >> 24) << 13) | (hi << 21);
                            const uint32_t recP = (lo >> 13) & (WP - 1);
                            uint32_t Y = have_pend ? pendY : recY, pid = have_pend ? pendP : recP;
                            bool retry = have_pend;
                            have_pend = false;
                            for (;;) {
                                const uint32_t slot = smem_atom_add(s_state + pid * 4, 1u) & 0xFFFFu;
                                if (slot < (uint32_t)WCAP) {
                                    smem_st(s_buf + (pid * WCAP + slot) * 4, Y);
                                    const uint32_t wr = smem_atom_add(s_state + pid * 4, 0x10000u) >> 16;
                                    if (wr == (uint32_t)WCAP - 1) flush_bin(pid);
                                } else if (retry) {
                                    KC_STAT(2);
                                    wide_fallback(wide_rec(Y, pid), 0x7Fu, table);
                                } else {
                                    KC_STAT(1);
                                    pendY = Y;
                                    pendP = pid;
                                    have_pend = true;
                                }
                                if (!retry) break;
                                retry = false;
                                Y = recY;
                                pid = recP;
                            }
                        }
                    }
                }
                cur16 = nxt;
                d += 512 % WA;
                if (d >= WA) d -= WA;
            }
            const uint32_t zero = kc_opaque_zero((uint32_t)d);  // pins the ring copies at the end of the iteration
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
    if (have_pend) wide_fallback(wide_rec(pendY, pendP), 0x7Fu, table);
    __syncthreads();
    for (int b = tid >> 5; b < WP; b += 32) {
        const uint32_t c = smem[b] & 0xFFFFu;
        const uint32_t base_off = b * region_cap;
        const uint32_t pos = smem[WP + b] - base_off;
        uint32_t stored = pos < region_cap ? pos : region_cap;
        if ((uint32_t)lane < c) {
            const uint32_t r = smem[2 * WP + b * WCAP + lane];
            if (pos + WCAP <= region_cap)
                my_slabs[base_off + pos + lane] = r;
            else
                wide_fallback(wide_rec(r, b), 0x7Fu, table);
        }
        if (pos + WCAP <= region_cap) stored = pos + c;
        if (lane == 0) counts[(uint64_t)b * gridDim.x + blockIdx.x] = stored;
    }
}

// ---------------------------------------------------------------------------
// Pass 1, second generation (KC_DENSE_PARTITION_WIDE2): the same 18-base records and slab layout as
// part_scatter7_kernel, rebuilt around what the ncu captures showed.  Round 1's kernel
// (profiles/r01_ncu_instruction_mix.txt): 339 warp instructions per 512-byte step, a quarter of them
// control flow; 2.2 % of the records in a divergent RED fallback that costs 22 % of the instructions;
// 2.56 x the necessary DRAM traffic.  Two earlier shapes of THIS kernel (profiles/r02_scatter7v2_history.txt):
// fully unrolled record slots = 13.6 K instructions, 36 % of the stalls on the instruction cache, 5.05 ms;
// rolled, warp-cooperative bin flush = issue-bound at 3227 instructions per super-step, 3.01 ms.
//
//   * SUPER-STEP = 7 x 512 bytes per warp = 3584 bases = exactly 512 records.  Every lane decodes its
//     seven coalesced 16-byte blocks, the 2-bit words go through a warp-private shared-memory tile
//     (one conflict-free store per block), and every lane reads its own 112 CONTIGUOUS bases back:
//     exactly 16 records per lane at bit offsets 14 r.  No per-lane alignment state and no idle record
//     slots (the kernels above run 4 or 3 slots per lane and step for 3.2 or 2.3 records).
//   * Three warp-uniform cases per super-step: no invalid base in it (every record is whole: no validity
//     arithmetic), every base invalid (inside an N run: nothing to do), mixed (the lane decodes the
//     bad-base masks of its own 128 bases again; rare).
//   * The state word of a bin is (generation << 16 | slots reserved): generation * 16 is also the bin's write
//     position in the CTA's private region, so the cursor array of the old kernel and its atomic are gone.  A
//     writer reserves a slot (returning atomic), stores, and bumps done[bin] (non-returning).  The lane that
//     takes slot 15 flushes the bin at once: it waits until done[bin] says that all 16 records of its generation
//     are stored, reads them with four 128-bit shared loads, opens the next generation with one store and
//     writes the 64-byte chunk with four 128-bit global stores.
//   * A record that meets a full bin WAITS IN A REGISTER for an extra record slot at the end of the
//     super-step.  Only a failure while another record already waits is counted with global REDs
//     (measured: < 0.5 % of the records; the old kernel sent 2.2 % there).
//
// No waiting on other lanes except the flusher's wait for done[], and the writers it waits for never wait.
// ---------------------------------------------------------------------------
constexpr int W2_BLOCKS = 7;            // 16-byte blocks per lane and super-step
constexpr int W2_TILE = 8 * 32;         // words of a warp's tile: 7 blocks + the next super-step's first one (halo)

#ifdef KC_EMU
#define KC_W2_PAUSE() emu::maybe_preempt_always()
#else
#define KC_W2_PAUSE() __nanosleep(20)
#endif

struct W2Stage {
    uint32_t s_state, s_done, s_bins;   // shared-window addresses
    uint32_t* my_slabs;                 // this CTA's regions
    uint32_t region_cap;
    uint32_t* table;
};

// region overflow (skewed input): the 16 records of a bin are counted directly.  Rare, slow, exact.
__device__ __noinline__ void w2_overflow(uint4 a, uint4 b, uint4 c, uint4 d, uint32_t pid, uint32_t* table) {
    KC_STAT(9);
    const uint32_t v[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll 1
    for (int q = 0; q < 16; q++) wide_fallback(wide_rec(v[q], pid), 0x7Fu, table);
}

// the lane that completed generation `gen` of bin `pid` writes it out
__device__ __forceinline__ void w2_flush(const W2Stage& c, uint32_t pid, uint32_t gen, uint32_t binaddr) {
    // done[pid] counts the records STORED into the bin since the launch: generation gen is complete at 16 (gen + 1)
    const uint32_t want = (gen + 1u) * WCAP;
    while (smem_ld(c.s_done + pid * 4) != want) {
        KC_STAT(8);  // a writer of this generation is between its atomic and its store
        // The pause is load-bearing: with an EMPTY retry path nvcc 12.9 compiled an earlier form of this wait to one
        // pass without the test (cuobjdump showed the loads, then the reset store), and ~1e-4 of the records were
        // flushed as zeros on a B200 while the sequentially consistent emulator saw nothing.
        KC_W2_PAUSE();
    }
    // (One bulk copy of the TMA unit instead of the eight LSU instructions below was measured too: scatter 2.03 ms
    // against 1.92 — the bin stays closed until the unit has read it.  profiles/r02_scatter7v2_history.txt.)
    const uint32_t pos = gen * WCAP;
    const uint4 v0 = smem_ld128(binaddr), v1 = smem_ld128(binaddr + 16), v2 = smem_ld128(binaddr + 32), v3 = smem_ld128(binaddr + 48);
    smem_st(c.s_state + pid * 4, (gen + 1u) << 16);  // the next generation is open
    if (pos + WCAP <= c.region_cap) {
        uint4* dst = reinterpret_cast<uint4*>(c.my_slabs + (pid * c.region_cap + pos));  // 64-byte aligned; < 2^32 words per CTA (host)
        dst[0] = v0;
        dst[1] = v1;
        dst[2] = v2;
        dst[3] = v3;
    } else {
        w2_overflow(v0, v1, v2, v3, pid, c.table);
    }
}

// one attempt: true = the record is in its bin.  Two shared atomics per record as in part_scatter_kernel, but the
// second one does not return (RED), nothing is ever reset but the state word, and the flusher is known from the FIRST
// atomic — it waits for done[] instead of being chosen by it.  (A first form of this kernel carried the generation's
// parity in bit 31 of every record and checked the 16 flags of a bin: one atomic per record, but the 16-way test was a
// quarter of the kernel's ALU-pipe instructions, and the ALU pipe is what bounds it: profiles/r02_scatter7v2_history.txt.)
__device__ __forceinline__ bool w2_stage(const W2Stage& c, uint32_t pid, uint32_t Y) {
    const uint32_t old = smem_atom_add(c.s_state + pid * 4, 1u);
    const uint32_t slot = old & 0xFFFFu;
    if (slot >= (uint32_t)WCAP) return false;
    const uint32_t binaddr = c.s_bins + pid * (WCAP * 4);
    smem_st(binaddr + slot * 4, Y);
    smem_red_add(c.s_done + pid * 4, 1u);  // after the store, in program order
    if (slot == (uint32_t)WCAP - 1) w2_flush(c, pid, old >> 16, binaddr);
    return true;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
part_scatter7v2_kernel(const uint4* __restrict__ base, uint64_t nsuper, uint32_t* __restrict__ table,
                       uint32_t* __restrict__ slabs, uint32_t* __restrict__ counts, uint32_t region_cap) {
    constexpr int NW = THREADS / 32;
    KC_DYN_SMEM(uint32_t, smem);
    // words: state[WP] | done[WP] | bins[WP][16] | tile[NW warps][256]
    W2Stage c;
    c.s_state = (uint32_t)__cvta_generic_to_shared(smem);
    c.s_done = c.s_state + WP * 4;
    c.s_bins = c.s_done + WP * 4;
    c.region_cap = region_cap;
    c.table = table;
    c.my_slabs = slabs + (uint64_t)blockIdx.x * WP * region_cap;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t s_tile = c.s_bins + WP * WCAP * 4 + (uint32_t)warp * (W2_TILE * 4);
    const uint32_t s_mine = s_tile + 7u * (uint32_t)lane * 4u;  // this lane's 8 words
    for (int b = tid; b < 2 * WP; b += THREADS) smem[b] = 0;  // generation 0, empty; nothing stored
    __syncthreads();

    const uint64_t nwarps = (uint64_t)gridDim.x * NW;
    const uint64_t w = (uint64_t)blockIdx.x * NW + warp;
    const uint64_t sb = w * nsuper / nwarps, se = (w + 1) * nsuper / nwarps;

    uint32_t pend_pid = 0, pend_Y = 0;
    bool have_pend = false;
    // a record that meets a full bin waits in a register for the extra slot at the end of the super-step; a second
    // one while the first still waits is counted with global REDs
    auto submit = [&](uint32_t pid, uint32_t Y) {
        if (!w2_stage(c, pid, Y)) {
            if (have_pend) {
                KC_STAT(2);
                wide_fallback(wide_rec(Y, pid), 0x7Fu, table);
            } else {
                KC_STAT(0);
                pend_pid = pid;
                pend_Y = Y;
                have_pend = true;
            }
        }
    };

    if (sb < se) {
        const uint4* ptr = base + sb * (W2_BLOCKS * 32) + lane;
        Decoded16 carry = kc_decode16(kc_ldg_stream(ptr));  // block 0 of the first super-step
        uint4 raw[W2_BLOCKS];
#pragma unroll
        for (int j = 0; j < W2_BLOCKS; j++) raw[j] = kc_ldg_stream(ptr + 32 * (j + 1));
        for (uint64_t s = sb; s < se; s++) {
            // decode blocks 1..7 of this super-step (block 7 = block 0 of the next one: halo now, carry later) and
            // refill every register with the block the NEXT iteration decodes: seven loads per lane stay in flight
            // for a whole super-step
            uint32_t anybad = carry.bad, allbad = carry.bad;
            smem_st(s_tile + lane * 4, carry.packed);
#pragma unroll
            for (int j = 0; j < W2_BLOCKS; j++) {
                const Decoded16 d = kc_decode16(raw[j]);
                raw[j] = kc_ldg_stream(ptr + 32 * (j + 1 + W2_BLOCKS));
                smem_st(s_tile + ((j + 1) * 32 + lane) * 4, d.packed);
                anybad |= d.bad;
                if (j < W2_BLOCKS - 1) allbad &= d.bad;
                if (j == W2_BLOCKS - 1) carry = d;
            }
            const uint4* const blk = ptr + 6 * lane;  // = base + s * 224 + 7 * lane: this lane's eight CONSECUTIVE blocks
            ptr += W2_BLOCKS * 32;
            const bool clean = !__any_sync(0xffffffffu, anybad != 0);
            // inside an N run: no window that starts in this super-step is valid
            if (!clean && __all_sync(0xffffffffu, allbad == 0xFFFFu)) continue;
            __syncwarp();
            uint64_t B01 = 0, B23 = 0;
            if (!clean) {
                KC_STAT(10);
                // validity of the 112 + 16 bases behind this lane's first base: the decoder runs again on the lane's
                // own region, this time for the masks (a mixed super-step is rare: ~1.5 % of the bench genome)
                uint32_t h[8];
#pragma unroll
                for (int q = 0; q < 8; q++) h[q] = kc_decode16(kc_ldg_stream(blk + q)).bad;
                B01 = (uint64_t)(h[0] | (h[1] << 16)) | ((uint64_t)(h[2] | (h[3] << 16)) << 32);
                B23 = (uint64_t)(h[4] | (h[5] << 16)) | ((uint64_t)(h[6] | (h[7] << 16)) << 32);
            }
            // The lane's 16 records start at bit 14 r of the 256-bit string tile[7 lane .. 7 lane + 7].  ONE rolled
            // loop body (two records) keeps the kernel small: three conflict-free shared loads per two records
            // instead of eight registers and 16 copies of the staging code.
#pragma unroll 1
            for (int i = 0; i < 8; i++) {
                const uint32_t bit = 28u * (uint32_t)i;  // warp-uniform
                const uint32_t wa = s_mine + (bit >> 5) * 4u;
                uint32_t sh = bit & 31u;
                uint32_t w0 = smem_ld(wa), w1 = smem_ld(wa + 4), w2 = smem_ld(wa + 8);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (sh >= 32u) {  // warp-uniform
                        sh -= 32u;
                        w0 = w1;
                        w1 = w2;
                        w2 = 0;  // (the second record never reaches past w2: 31 + 14 + 36 <= 96)
                    }
                    const uint32_t lo = __funnelshift_r(w0, w1, sh);   // bases 0 .. 15
                    const uint32_t hiw = __funnelshift_r(w1, w2, sh);  // bases 16, 17 in bits [0, 4)
                    const uint32_t hi = hiw & 0xFu;
                    bool valid = true;
                    if (!clean) {
                        // 18 bad bits at bit 7r of B, then "any bad base among 12" per window start
                        const uint32_t o = 7u * (uint32_t)(2 * i + h);  // <= 105
                        const uint64_t cur = (o & 64u) ? B23 : B01, nxt = (o & 64u) ? 0ull : B23;
                        const uint32_t os = o & 63u;
                        uint32_t bb = (uint32_t)(cur >> os);
                        if (os > 46u) bb |= (uint32_t)(nxt << (64u - os));
                        bb &= 0x3FFFFu;
                        const uint32_t y2 = bb | (bb >> 1), y4 = y2 | (y2 >> 2), y8 = y4 | (y4 >> 4);
                        const uint32_t okr = ~(y8 | (y4 >> 8)) & 0x7Fu;
                        valid = okr == 0x7Fu;
                        if (!valid && okr) wide_fallback((uint64_t)lo | ((uint64_t)hi << 32), okr, table);  // next to an invalid base
                    }
                    // Y = record bits [0,13) | bits [24,36) << 13, built with two funnel shifts; its bits >= 25 are
                    // whatever follows the record (pass 2 masks every index it takes from Y)
                    const uint32_t t12 = __funnelshift_r(lo, hiw, 24);
                    if (valid) submit((lo << 8) >> 21, __funnelshift_r(lo << 19, t12, 19));
                    sh += 14u;
                }
            }
            if (__any_sync(0xffffffffu, have_pend)) {  // a slot for the records that wait
                if (have_pend) {
                    have_pend = false;
                    submit(pend_pid, pend_Y);  // (a failure puts it back)
                }
            }
            __syncwarp();  // every lane has read its words: the tile may be overwritten
        }
    }
    if (have_pend) {  // (only after an all-N tail)
        if (!w2_stage(c, pend_pid, pend_Y)) wide_fallback(wide_rec(pend_Y, pend_pid), 0x7Fu, table);
    }
    // final flush of the partially filled bins, then publish the region lengths
    __syncthreads();
    for (int b = warp; b < WP; b += NW) {
        const uint32_t st = smem[b];
        uint32_t n = st & 0xFFFFu;  // < WCAP: a full bin was flushed by the lane that filled it
        if (n > (uint32_t)WCAP) n = WCAP;
        const uint32_t pos = (st >> 16) * WCAP;
        const uint32_t base_off = b * region_cap;
        uint32_t stored = pos < region_cap ? pos : region_cap;
        if ((uint32_t)lane < n) {
            const uint32_t r = smem[2 * WP + b * WCAP + lane];
            if (pos + WCAP <= region_cap)
                c.my_slabs[base_off + pos + lane] = r;
            else
                wide_fallback(wide_rec(r, b), 0x7Fu, table);
        }
        if (pos + WCAP <= region_cap) stored = pos + n;
        if (lane == 0) counts[(uint64_t)b * gridDim.x + blockIdx.x] = stored;
    }
}

__device__ __forceinline__ uint32_t nibsum(uint32_t v) {  // sum of the eight 4-bit fields
    const uint32_t a = (v & 0x0F0F0F0Fu) + ((v >> 4) & 0x0F0F0F0Fu);  // bytes <= 30
    return (a * 0x01010101u) >> 24;                                    // <= 120
}
__device__ __forceinline__ uint32_t nibsum_par(uint32_t v, uint32_t odd) {  // the four fields of one parity
    const uint32_t a = (v >> (4 * odd)) & 0x0F0F0F0Fu;
    return (a * 0x01010101u) >> 24;
}

// Pass 2 for the records of part_scatter7_kernel.
__global__ void __launch_bounds__(1024, 1)
part_count7_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slabs, const uint32_t* __restrict__ counts,
                   uint32_t region_cap, uint32_t nregions, uint32_t* __restrict__ work_counter) {
    KC_DYN_SMEM(uint32_t, bins);  // 57344 words: the recount needs 7 x 8192; the nibble tables use 36864
    __shared__ uint32_t s_part, s_bad;
    __shared__ unsigned long long s_red[128];
    const uint32_t s_bins = (uint32_t)__cvta_generic_to_shared(bins);
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t W3 = 16384, W6 = 32768, NW = 36864;  // word offsets of T3, T6; words of the three tables
    for (;;) {
        if (tid == 0) s_part = atomicAdd(work_counter, 1u);
        {
            uint4* b4 = reinterpret_cast<uint4*>(bins);
            for (int i = tid; i < (int)NW / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= (uint32_t)WP) break;
        uint32_t nrec = 0;
        auto count_rec = [&](uint32_t Y) {
            const uint32_t i0 = Y & 0x1FFFFu, i3 = (Y >> 6) & 0x1FFFFu, i6 = Y >> 12;
            smem_red_add(s_bins + (i0 & 0x3FFFu) * 4, 1u << (4 * (i0 >> 14)));
            smem_red_add(s_bins + (W3 + (i3 & 0x3FFFu)) * 4, 1u << (4 * (i3 >> 14)));
            smem_red_add(s_bins + (W6 + (i6 & 0xFFFu)) * 4, (i6 & 0x1000u) ? 0x10000u : 1u);
            nrec++;
        };
        auto count_rec_classic = [&](uint32_t Y) {  // seven 32-bit sub-tables: window r's bin = Y bits [2r, 2r+13)
#pragma unroll
            for (int r = 0; r < WA; r++) smem_red_add(s_bins + (r * 8192 + ((Y >> (2 * r)) & 0x1FFFu)) * 4, 1u);
        };
        auto for_each_record = [&](auto&& f) {
            for (uint32_t reg = warp; reg < nregions; reg += 32) {
                const uint32_t n = counts[(uint64_t)part * nregions + reg];
                const uint32_t* src = slabs + ((uint64_t)reg * WP + part) * region_cap;
                const uint4* src4 = reinterpret_cast<const uint4*>(src);
                const uint32_t n4 = n >> 2;
                uint32_t i = lane;
                for (; i + 32 < n4; i += 64) {
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
        uint32_t a0 = 0, a3 = 0, a6 = 0;  // per-warp sums fit 32 bits; one REDUX each
        for (int i = tid; i < 16384; i += 1024) {
            a0 += nibsum(bins[i]);
            a3 += nibsum(bins[W3 + i]);
        }
        for (int i = tid; i < 4096; i += 1024) {
            const uint32_t v = bins[W6 + i];
            a6 += (v & 0xFFFFu) + (v >> 16);
        }
        const unsigned long long s0 = __reduce_add_sync(0xffffffffu, a0), s3 = __reduce_add_sync(0xffffffffu, a3),
                                 s6 = __reduce_add_sync(0xffffffffu, a6), asum = __reduce_add_sync(0xffffffffu, nrec);
        if (lane == 0) {
            s_red[warp] = s0;
            s_red[32 + warp] = s3;
            s_red[64 + warp] = s6;
            s_red[96 + warp] = asum;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long a = 0, b = 0, c = 0, n = 0;
            for (int q = 0; q < 32; q++) {
                a += s_red[q];
                b += s_red[32 + q];
                c += s_red[64 + q];
                n += s_red[96 + q];
            }
            s_bad = (a == n && b == n && c == n) ? 0u : 1u;
#ifdef KC_EMU  // the emulator tests also force the recount (every odd partition)
            if (getenv("KC_EMU_FORCE_PAIR_RECOUNT") && (part & 1u)) s_bad = 1u;
#endif
        }
        __syncthreads();
        if (s_bad) {  // a field wrapped: recount this partition with 32-bit bins
            KC_STAT(7);
            {
                uint4* b4 = reinterpret_cast<uint4*>(bins);
                for (int i = tid; i < 57344 / 4; i += 1024) b4[i] = make_uint4(0, 0, 0, 0);
            }
            __syncthreads();
            for_each_record(count_rec_classic);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < WA; r++) {
                const int lowbits = 13 - 2 * r;  // 13, 11, ..., 1
                for (int f = tid; f < 8192; f += 1024) {
                    const uint32_t v = bins[r * 8192 + f];
                    if (v) {
                        const uint32_t low = (uint32_t)f & ((1u << lowbits) - 1u), high = (uint32_t)f >> lowbits;
                        global_red_add(table + (low | (part << lowbits) | (high << (24 - 2 * r))), v);
                    }
                }
            }
        } else {
            // Fold and flush.  A 4-bit field f of T0/T3 lives in word f & 0x3FFF, nibble f >> 14.
            // T0: f = low13 | hi4 << 13  (low13 = record bits [0,13), hi4 = bits [24,28))
            //   window 0 = low13 | key << 13                           sum over hi4
            //   window 1 = low13 >> 2 | key << 11 | (hi4 & 3) << 22    sum over low13 & 3 and hi4 >> 2
            //   window 2 = low13 >> 4 | key << 9 | hi4 << 20           sum over low13 & 15
            // T3: f = low7 | hi10 << 7  (low7 = record bits [6,13), hi10 = bits [24,34))
            //   window 3 = low7 | key << 7 | (hi10 & 63) << 18         sum over hi10 >> 6
            //   window 4 = low7 >> 2 | key << 5 | (hi10 & 255) << 16   sum over low7 & 3 and hi10 >> 8
            //   window 5 = low7 >> 4 | key << 3 | hi10 << 14           sum over low7 & 15
            // T6: f = bit12 | hi12 << 1:  window 6 = bit12 | key << 1 | hi12 << 12
            for (int x = tid; x < 8192; x += 1024) {
                const uint32_t v0 = nibsum(bins[x]) + nibsum(bins[x + 8192]);
                if (v0) global_red_add(table + ((uint32_t)x | (part << 13)), v0);
                const uint32_t v3 = nibsum(bins[W3 + x]) + nibsum(bins[W3 + x + 8192]);
                if (v3) global_red_add(table + (((uint32_t)x & 127u) | (part << 7) | (((uint32_t)x >> 7) << 18)), v3);
            }
            for (int y = tid; y < 4096; y += 1024) {
                // a block of 4 words = fields with the same low bits >> 2; nibble parity = bit 14 of f
                const uint4 w = *reinterpret_cast<const uint4*>(bins + 4 * y);
                const uint4 u = *reinterpret_cast<const uint4*>(bins + W3 + 4 * y);
#pragma unroll
                for (uint32_t odd = 0; odd < 2; odd++) {
                    // T0: word bit 13 = hi4 & 1, nibble = hi4 >> 1: (hi4 & 3) = (y >> 11) | odd << 1
                    const uint32_t v1 = nibsum_par(w.x, odd) + nibsum_par(w.y, odd) + nibsum_par(w.z, odd) + nibsum_par(w.w, odd);
                    if (v1) global_red_add(table + (((uint32_t)y & 2047u) | (part << 11) | ((((uint32_t)y >> 11) | (odd << 1)) << 22)), v1);
                    // T3: word bits [7,14) = hi10 & 127, nibble = hi10 >> 7: (hi10 & 255) = (y >> 5) | odd << 7
                    const uint32_t v4 = nibsum_par(u.x, odd) + nibsum_par(u.y, odd) + nibsum_par(u.z, odd) + nibsum_par(u.w, odd);
                    if (v4) global_red_add(table + (((uint32_t)y & 31u) | (part << 5) | ((((uint32_t)y >> 5) | (odd << 7)) << 16)), v4);
                }
            }
            for (int z = tid; z < 1024; z += 1024) {
                // 16 consecutive words: one bin per nibble position
                uint32_t acc0[8], acc3[8];
#pragma unroll
                for (int nb = 0; nb < 8; nb++) acc0[nb] = acc3[nb] = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 w = *reinterpret_cast<const uint4*>(bins + 16 * z + 4 * q);
                    const uint4 u = *reinterpret_cast<const uint4*>(bins + W3 + 16 * z + 4 * q);
#pragma unroll
                    for (int nb = 0; nb < 8; nb++) {
                        acc0[nb] += ((w.x >> (4 * nb)) & 15u) + ((w.y >> (4 * nb)) & 15u) + ((w.z >> (4 * nb)) & 15u) + ((w.w >> (4 * nb)) & 15u);
                        acc3[nb] += ((u.x >> (4 * nb)) & 15u) + ((u.y >> (4 * nb)) & 15u) + ((u.z >> (4 * nb)) & 15u) + ((u.w >> (4 * nb)) & 15u);
                    }
                }
                // T0: z = low13 >> 4 (9 bits) | (hi4 & 1) << 9; nibble = hi4 >> 1
                // T3: z = low7 >> 4 (3 bits) | (hi10 & 127) << 3; nibble = hi10 >> 7
#pragma unroll
                for (uint32_t nb = 0; nb < 8; nb++) {
                    if (acc0[nb]) global_red_add(table + (((uint32_t)z & 511u) | (part << 9) | ((((uint32_t)z >> 9) | (nb << 1)) << 20)), acc0[nb]);
                    if (acc3[nb]) global_red_add(table + (((uint32_t)z & 7u) | (part << 3) | ((((uint32_t)z >> 3) | (nb << 7)) << 14)), acc3[nb]);
                }
            }
            for (int f = tid; f < 8192; f += 1024) {
                const uint32_t wv = bins[W6 + (f & 0xFFF)];
                const uint32_t v6 = (f & 0x1000) ? (wv >> 16) : (wv & 0xFFFFu);
                if (v6) global_red_add(table + (((uint32_t)f & 1u) | (part << 1) | (((uint32_t)f >> 1) << 12)), v6);
            }
        }
        __syncthreads();
    }
}

}  // namespace

// Host side: same geometry as dense_partition<> (dense.cu) with A = 7.
int kc_dense_partition_wide(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                            uint32_t* d_table, cudaStream_t st) {
    constexpr int DEPTH = 3;
    const ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, WK);
    const uint64_t G0 = (max(g.lo, g.wlo) + 511) >> 9;
    const uint64_t lim = min(g.hi, g.whi);
    const uint64_t Gl = lim >> 9;
    const uint64_t G1 = Gl > (uint64_t)(2 * DEPTH + 2) ? Gl - (2 * DEPTH + 2) : 0;
    if (G1 <= G0 + 64) return kc_dense_direct_range(ctx, d_data, nbytes, win_begin, win_end, WK, d_table, st);
    const uint64_t ngroups = (G1 - G0) / DEPTH * DEPTH;
    const uint64_t nrec = (ngroups * 512 + WA - 1) / WA;  // records start at (G0<<9) + 7 m
    const uint64_t shift = g.lo;
    const uint64_t head_end = (G0 << 9) - shift;
    const uint64_t tail_begin = (G0 << 9) + nrec * WA - shift;
    const uint64_t want = (ngroups + 31) / 32;
    const int grid1 = (int)(want > (uint64_t)ctx->sm_count ? (uint64_t)ctx->sm_count : want);
    uint64_t cap = nrec / ((uint64_t)WP * grid1);
    cap = cap + cap / 8 + 4 * WCAP;
    cap = (cap + 31) / 32 * 32;
    if ((uint64_t)WP * cap >= (1ull << 32)) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition region too large");
    const size_t nregions = (size_t)WP * grid1;
    const size_t ctl_pad = ((nregions + 64) * sizeof(uint32_t) + 255) & ~(size_t)255;
    int rc = kc_scratch_reserve(ctx, ctl_pad + nregions * cap * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t* counts = (uint32_t*)ctx->scratch;
    uint32_t* work = counts + nregions;
    uint32_t* slabs = (uint32_t*)((char*)ctx->scratch + ctl_pad);
    KC_CUDA(ctx, cudaMemsetAsync(work, 0, 64 * sizeof(uint32_t), st));
    const size_t smem1 = (size_t)(2 * WP + WP * WCAP) * sizeof(uint32_t);
    const size_t smem2 = 57344 * sizeof(uint32_t);
    const uint4* base = g.abase + (G0 << 5);
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
    KC_CUDA(ctx, cudaFuncSetAttribute(part_scatter7_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    KC_LAUNCH(part_scatter7_kernel<DEPTH>, grid1, 1024, smem1, st, base, ngroups, d_table, slabs, counts, (uint32_t)cap);
    KC_LAUNCH_CHECK(ctx, "part_scatter7_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
    const int grid2 = ctx->sm_count < WP ? ctx->sm_count : WP;
    KC_CUDA(ctx, cudaFuncSetAttribute(part_count7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    KC_LAUNCH(part_count7_kernel, grid2, 1024, smem2, st, d_table, slabs, counts, (uint32_t)cap, (uint32_t)grid1, work);
    KC_LAUNCH_CHECK(ctx, "part_count7_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[2], st));
    const bool timing = ctx->timing;
    ctx->timing = false;
    rc = KC_OK;
    if (head_end > win_begin) rc = kc_dense_direct_range(ctx, d_data, nbytes, win_begin, head_end, WK, d_table, st);
    if (!rc && tail_begin < win_end) rc = kc_dense_direct_range(ctx, d_data, nbytes, tail_begin, win_end, WK, d_table, st);
    ctx->timing = timing;
    if (timing) ctx->timed_kernels = 2;
    return rc;
}

// Host side of KC_DENSE_PARTITION_WIDE2: part_scatter7v2_kernel + part_count7_kernel.  The interior is a whole
// number of super-steps (7 groups of 512 bytes = 512 records), so the records tile it exactly.
int kc_dense_partition_wide2(kc_ctx* ctx, const char* d_data, uint64_t nbytes, uint64_t win_begin, uint64_t win_end,
                             uint32_t* d_table, cudaStream_t st) {
    const ScanGeom g = kc_make_geom(d_data, nbytes, win_begin, win_end, WK);
    const uint64_t G0 = (max(g.lo, g.wlo) + 511) >> 9;
    const uint64_t lim = min(g.hi, g.whi);
    const uint64_t Gl = lim >> 9;
    // the last warp prefetches one super-step + one block past its last super-step: 2 * 7 + 2 groups stay readable
    const uint64_t G1 = Gl > (uint64_t)(2 * W2_BLOCKS + 2) ? Gl - (2 * W2_BLOCKS + 2) : 0;
    if (G1 <= G0 + 64) return kc_dense_direct_range(ctx, d_data, nbytes, win_begin, win_end, WK, d_table, st);
    const uint64_t nsuper = (G1 - G0) / W2_BLOCKS;
    const uint64_t ngroups = nsuper * W2_BLOCKS;
    const uint64_t nrec = nsuper * 512;  // records start at (G0<<9) + 7 m and tile the interior groups exactly
    const uint64_t shift = g.lo;
    const uint64_t head_end = (G0 << 9) - shift;
    const uint64_t tail_begin = ((G0 + ngroups) << 9) - shift;
    const uint64_t want = (nsuper + 31) / 32;
    const int grid1 = (int)(want > (uint64_t)ctx->sm_count ? (uint64_t)ctx->sm_count : want);
    uint64_t cap = nrec / ((uint64_t)WP * grid1);
    cap = cap + cap / 8 + 4 * WCAP;
    cap = (cap + 31) / 32 * 32;
    if ((uint64_t)WP * cap >= (1ull << 32) || cap / WCAP >= 65535)  // the bin generation (= chunks written) is a 16-bit field
        return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "partition region too large");
    const size_t nregions = (size_t)WP * grid1;
    const size_t ctl_pad = ((nregions + 64) * sizeof(uint32_t) + 255) & ~(size_t)255;
    int rc = kc_scratch_reserve(ctx, ctl_pad + nregions * cap * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t* counts = (uint32_t*)ctx->scratch;
    uint32_t* work = counts + nregions;
    uint32_t* slabs = (uint32_t*)((char*)ctx->scratch + ctl_pad);
    KC_CUDA(ctx, cudaMemsetAsync(work, 0, 64 * sizeof(uint32_t), st));
    // 1024 threads: the 512-thread build (126 registers, no spills) measured 8 % slower on B200
    const size_t smem1 = (size_t)(2 * WP + WP * WCAP + 32 * W2_TILE) * sizeof(uint32_t);
    const size_t smem2 = 57344 * sizeof(uint32_t);
    const uint4* base = g.abase + (G0 << 5);
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[0], st));
    KC_CUDA(ctx, cudaFuncSetAttribute(part_scatter7v2_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    KC_LAUNCH(part_scatter7v2_kernel<1024>, grid1, 1024, smem1, st, base, nsuper, d_table, slabs, counts, (uint32_t)cap);
    KC_LAUNCH_CHECK(ctx, "part_scatter7v2_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[1], st));
    const int grid2 = ctx->sm_count < WP ? ctx->sm_count : WP;
    KC_CUDA(ctx, cudaFuncSetAttribute(part_count7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    KC_LAUNCH(part_count7_kernel, grid2, 1024, smem2, st, d_table, slabs, counts, (uint32_t)cap, (uint32_t)grid1, work);
    KC_LAUNCH_CHECK(ctx, "part_count7_kernel");
    if (ctx->timing) KC_CUDA(ctx, cudaEventRecord(ctx->tev[2], st));
    const bool timing = ctx->timing;
    ctx->timing = false;
    rc = KC_OK;
    if (head_end > win_begin) rc = kc_dense_direct_range(ctx, d_data, nbytes, win_begin, head_end, WK, d_table, st);
    if (!rc && tail_begin < win_end) rc = kc_dense_direct_range(ctx, d_data, nbytes, tail_begin, win_end, WK, d_table, st);
    ctx->timing = timing;
    if (timing) ctx->timed_kernels = 2;
    return rc;
}

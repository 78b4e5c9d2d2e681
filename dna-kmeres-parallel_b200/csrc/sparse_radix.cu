// sparse_radix.cu — KC_SPARSE_RADIX: sparse k-mer counting (13 <= k <= 31) as an MSD radix
// partition whose leaves are sorted in shared memory.
//
// Why: the hash path (sparse.cu) does one random DRAM access per window into a multi-GB
// table (config 4 at full scale: 13 G inserts, 1.77 s); the sort path pays 6+ device-wide
// radix passes over 64-bit codes.  Here every global access is a coalesced stream:
//
//   K1 sp_scatter_kernel   scan ASCII -> 2k-bit LE code per valid window; partition by the top
//                          KB1 bits of the code; the REMAINING bits (a 32- or 64-bit record)
//                          are staged per partition in shared memory and leave the SM in
//                          64-byte chunks into this CTA's private region of the partition.
//   K2 sp_leaf_kernel      one level-1 partition per CTA at a time: (A) the same staged
//                          scatter by the next KB2 bits into CTA-private scratch; (B) each
//                          of the 2^KB2 leaves (a few thousand records) is sorted in shared
//                          memory — counting sort on the top <=10 remaining bits, one
//                          sub-bucket per thread finished by insertion sort — run-length
//                          encoded, and its (code,count) runs appended to a temporary list.
//   K3 sp_scan_kernel      exclusive scan of the leaves' distinct counts (leaf order = code order)
//   K4 sp_gather_kernel    leaf runs -> final arrays.  The result is sorted by code because
//                          partition, leaf, sub-bucket and in-bucket order all follow the
//                          code's most significant bits: no device-wide sort anywhere.
//
// Exactness: a record is dropped only when a private region, a leaf buffer or the
// temporary list overflows, or when both staging bins of a partition stay full for
// SP_MAX_TRIES attempts; each of these raises the `failed` flag and the caller
// (kc_count_sparse) then recounts with the hash path.  Uniform data never gets there;
// heavily skewed data (poly-A) does, by design — see DESIGN.md.
//
// Window semantics are the reference's (main.cu:636-646): k consecutive bytes, counted iff
// all are upper-case ACGT; code = sum code(s[p]) * 4^p (utils.h:30-47).
#include "common.cuh"

#ifdef KC_EMU
#define KC_SPIN_PAUSE() emu::maybe_preempt_always()
#else
#define KC_SPIN_PAUSE() __nanosleep(32)
#endif

namespace {

constexpr int SP_THREADS = 1024;
constexpr int SP_MAX_TRIES = 256;
// why a run gave up (bits of SpCtl::failed, reported in the error text)
enum { SP_FAIL_REGION = 1, SP_FAIL_STAGING = 2, SP_FAIL_LEAF = 4, SP_FAIL_RUNLIST = 8 };

// ---------------------------------------------------------------------------
// Staged scatter of fixed-size records into P private regions (shared by K1 and K2).
//
// Shared memory: state[P][2] | cur[P] | buf[P][2][CAP].  Every partition has TWO bins of
// CAP records (one 64-byte chunk each).  state word: low 24 bits = slots reserved, high 8
// bits = slots written.
//   writer : slot = atom.add(state, 1) & 0xFFFFFF
//            slot <  CAP: store the record; w = atom.add(state, 1<<24) >> 24; the writer that
//                         makes w == CAP-1 has seen every slot written and flushes the bin
//                         by itself (4 x 128-bit shared loads, state = 0, 4 x 128-bit global
//                         stores to cur[p], which it advances by CAP);
//            slot >= CAP: the bin is full and its flush is in flight: take the partition's
//                         other bin; if that one is full as well, pause and try again.
// Shared-memory requests of one thread are performed in program order, so a writer's
// store precedes its "written" increment, and the flusher's loads precede its reset —
// the same protocol the dense partition path runs on the GPU (dense.cu), plus the second
// bin, which replaces that path's "count it directly with global REDs" fallback (a
// sparse result has no table to RED into).  Failed attempts inflate only the reserved
// field (24 bits: SP_MAX_TRIES x 1024 threads cannot carry into the written field), and the
// reset wipes them.  (Round 2 also ran the form of dense_wide.cu here — the written count as a
// red.shared without return value and the writer of the LAST slot waiting for it: scatter 82.6 vs
// 81.9 ms, leaf kernel 152.3 vs 150.3 ms at config 4; the second round trip is not what these
// kernels wait for — they are issue-bound at 4.4 warp instructions per window — so the form
// without a wait loop stays.)
// ---------------------------------------------------------------------------
template <typename RecT, int P_>
struct Stager {
    static constexpr int P = P_;
    static constexpr int CAP = 64 / (int)sizeof(RecT);  // records per 64-byte chunk
    static constexpr size_t SMEM_BYTES = (size_t)P * 2 * 4 + (size_t)P * 4 + (size_t)P * 2 * 64;

    uint32_t s_state, s_cur, s_buf;  // shared-window addresses
    RecT* region;                    // this CTA's region of partition p: region[p * pstride + i], i < cap
    uint64_t pstride;                // records between the regions of consecutive partitions
    uint32_t cap;                    // records per (CTA, partition) region; multiple of CAP
    uint32_t* failed;

    __device__ __forceinline__ void init(uint32_t* smem_words, RecT* region_, uint64_t pstride_, uint32_t cap_,
                                         uint32_t* failed_) {
        s_state = (uint32_t)__cvta_generic_to_shared(smem_words);
        s_cur = s_state + P * 2 * 4;
        s_buf = s_cur + P * 4;
        region = region_;
        pstride = pstride_;
        cap = cap_;
        failed = failed_;
        for (int b = threadIdx.x; b < P; b += blockDim.x) {
            smem_words[2 * b] = 0;
            smem_words[2 * b + 1] = 0;
            smem_words[2 * P + b] = 0;  // cur[p]: records of partition p already written by this CTA
        }
    }

    __device__ __forceinline__ void flush_bin(uint32_t bin /* = 2*p + x */) {
        const uint32_t src = s_buf + bin * 64;
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) v[q] = smem_ld128(src + 16 * q);
        smem_st(s_state + bin * 4, 0u);  // every slot has been read: the bin is free again
        const uint32_t p = bin >> 1;
        const uint32_t pos = smem_atom_add(s_cur + p * 4, (uint32_t)CAP);
        if (pos + CAP <= cap) {
            uint4* dst = reinterpret_cast<uint4*>(region + (uint64_t)p * pstride + pos);  // 64-byte aligned
#pragma unroll
            for (int q = 0; q < 4; q++) dst[q] = v[q];  // (two 256-bit stores measured slower: profiles/r02_microbench3.txt)
        } else {
            atomicOr(failed, (uint32_t)SP_FAIL_REGION);  // region full (skewed input): the caller recounts with the hash path
        }
    }

    __device__ __forceinline__ void stage(uint32_t p, RecT rec, uint32_t pref) {
        uint32_t bin = 2 * p + (pref & 1u);
        for (int tries = 0; tries < SP_MAX_TRIES; tries++) {
            const uint32_t sa = s_state + bin * 4;
            const uint32_t slot = smem_atom_add(sa, 1u) & 0xFFFFFFu;
            if (slot < (uint32_t)CAP) {
                if (sizeof(RecT) == 8)
                    smem_st64(s_buf + bin * 64 + slot * 8, (uint64_t)rec);
                else
                    smem_st(s_buf + bin * 64 + slot * 4, (uint32_t)rec);
                const uint32_t wr = smem_atom_add(sa, 1u << 24) >> 24;
                if (wr == (uint32_t)CAP - 1) flush_bin(bin);
                return;
            }
            bin ^= 1u;
            KC_STAT(3);
            if (tries & 1) {
                KC_STAT(4);
                KC_SPIN_PAUSE();  // both bins were full
            }
        }
        KC_STAT(5);
        atomicOr(failed, (uint32_t)SP_FAIL_STAGING);
    }

    // after a CTA barrier: write the partially filled bins and leave the region's record
    // count in cur[p] (shared) — counts_out[p * stride + col] too when counts_out != nullptr
    __device__ __forceinline__ void finish(uint32_t* smem_words, uint32_t* counts_out, uint32_t stride, uint32_t col) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const RecT* bufp = reinterpret_cast<const RecT*>(smem_words + 3 * P);
        for (int p = warp; p < P; p += nw) {
            const uint64_t base_off = (uint64_t)p * pstride;
            uint32_t pos = smem_words[2 * P + p];  // chunks written so far (may exceed cap after a failure)
#pragma unroll
            for (int x = 0; x < 2; x++) {
                uint32_t c = smem_words[2 * p + x] & 0xFFFFFFu;  // < CAP: full bins were flushed by their last writer
                if (c > (uint32_t)CAP) c = CAP;                  // (only after a give-up, which has raised `failed`)
                if (c > 0 && pos <= cap) {
                    if (pos + c <= cap) {
                        if ((uint32_t)lane < c) region[base_off + pos + lane] = bufp[(2 * p + x) * CAP + lane];
                    } else if (lane == 0) {
                        atomicOr(failed, (uint32_t)SP_FAIL_REGION);
                    }
                }
                pos += c;
            }
            __syncwarp();
            if (lane == 0) {
                const uint32_t stored = pos < cap ? pos : cap;
                smem_words[2 * P + p] = stored;
                if (counts_out) counts_out[(uint64_t)p * stride + col] = stored;
            }
        }
    }
};

template <int KB1_, int KB2_, int LEAF_CAP_>
struct SpShape {
    static constexpr int KB1 = KB1_, KB2 = KB2_;
    static constexpr int P1 = 1 << KB1_, P2 = 1 << KB2_;
    static constexpr int LEAF_CAP = LEAF_CAP_;  // leaf size unit: a leaf may hold LEAF_CAP / 2 distinct 32-bit records (half for 64-bit)
};

struct SpCtl {  // one 64-byte control block in device memory
    uint32_t work;               // K2 partition queue
    uint32_t failed;             // any overflow: the result is invalid
    unsigned long long out_cursor;  // entries appended to the temporary run list
    unsigned long long total;       // K3: sum of the leaves' distinct counts
    uint32_t pad[10];
};

// ---------------------------------------------------------------------------
// K1: ASCII -> level-1 partitions
// ---------------------------------------------------------------------------
template <typename Shape, typename R1T, int HALO>
__global__ void __launch_bounds__(SP_THREADS, 1)
sp_scatter_kernel(ScanGeom g, R1T* __restrict__ slabs1, uint32_t* __restrict__ counts1, uint32_t cap1, SpCtl* ctl,
                  int rbits, uint32_t round) {
    KC_DYN_SMEM(uint32_t, smem);
    using St = Stager<R1T, Shape::P1>;
    St st;
    // slabs are partition-major, slabs1[p][cta][cap1]: the regions of a RANGE of partitions are one
    // contiguous block, which is what the multi-GPU path sends to the rank that owns the range
    st.init(smem, slabs1 + (uint64_t)blockIdx.x * cap1, (uint64_t)gridDim.x * cap1, cap1, &ctl->failed);
    __syncthreads();
    const int k = g.k;
    // ROUNDS: the top `rbits` bits of the code select the round a window belongs to; this launch keeps the windows of
    // `round` only, and everything below works on the remaining cb = 2k - rbits bits (see kc_sparse_radix_plan)
    const int cb = 2 * k - rbits;
    const int r1bits = cb - Shape::KB1;
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;
    const uint64_t r1mask = (1ull << r1bits) - 1ull;
    const uint32_t pref = threadIdx.x >> 5;

    // groups are dealt out evenly (warp w: [w n / W, (w+1) n / W)): the private regions are sized
    // for an even split, and a ceil-sized share would leave the last CTA idle on small inputs
    const uint64_t ngroups = g.g_end - g.g_begin;
    const uint64_t nwarps = (uint64_t)gridDim.x * (SP_THREADS / 32);
    const uint64_t w = (uint64_t)blockIdx.x * (SP_THREADS / 32) + (threadIdx.x >> 5);
    const uint64_t gb = g.g_begin + w * ngroups / nwarps;
    const uint64_t ge = g.g_begin + (w + 1) * ngroups / nwarps;
    kc_warp_scan<HALO>(g, gb, ge, [&](const LaneWindow<HALO>& lw, uint64_t) {
        // once anything has overflowed the result is void: stop staging, so that a hopelessly
        // skewed input (every record into one partition) fails in microseconds, not minutes
        if (*(volatile uint32_t*)&ctl->failed) return;
        // a loop over the set bits, not 16 unrolled copies: the staging code is long and 16
        // copies of it are 56 KB of SASS, more than the instruction cache holds
        uint32_t m = lw.ok & 0xFFFFu;
        if (rbits >= 2) {
            // ROUNDS: keep the windows whose top rbits code bits are `round` — for all 16 windows of the lane at once,
            // before any code is built (a per-window test cost 15 instructions for each of the 7 in 8 windows that
            // belong to other rounds: 138 -> 102 ms per round of config 5, which has 8 on one GPU; with two rounds
            // the per-window test below is the cheaper one, 83 vs 89 ms).  The code of window j is the bit
            // range [2j, 2j + 2k) of the lane's 2-bit stream p0|p1|p2, its top bits the range [2j + cb, 2j + 2k):
            // S = stream >> cb holds them at stride 2, and plane b of all 16 windows is compared in one XOR.
            const uint64_t lo64 = (uint64_t)lw.p0 | ((uint64_t)lw.p1 << 32);
            const uint64_t S = (lo64 >> cb) | ((uint64_t)lw.p2 << (64 - cb));  // 22 - 8 <= cb <= 62: both shifts defined
            uint32_t differs = 0;
            for (int b = 0; b < rbits; b++) differs |= (uint32_t)(S >> b) ^ (0u - ((round >> b) & 1u));
            uint32_t x = ~differs & 0x55555555u;  // bit 2j: window j is of this round
            x = (x | (x >> 1)) & 0x33333333u;
            x = (x | (x >> 2)) & 0x0F0F0F0Fu;
            x = (x | (x >> 4)) & 0x00FF00FFu;
            x = (x | (x >> 8)) & 0x0000FFFFu;
            m &= x;
        }
        while (m) {
            const int j = __ffs((int)m) - 1;
            m &= m - 1u;
            const uint64_t code = lw.code64(j, kmask);
            if (rbits == 1 && (uint32_t)(code >> cb) != round) continue;
            st.stage((uint32_t)(code >> r1bits) & (Shape::P1 - 1), (R1T)(code & r1mask), pref + (uint32_t)j);
        }
    });
    __syncthreads();
    st.finish(smem, counts1, gridDim.x, blockIdx.x);
}

// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = sum.
// ONE barrier: every warp scans the 32 warp totals itself (round 2's capture of sp_leaf_kernel had 29 % of its stall
// samples at the second barrier of the old form, 31 warps waiting for warp 0).  `s_warp` = 2 x 32 words used
// alternately: `calls` is the caller's running count of scans (the same in every thread), so a warp that is still
// reading the totals of scan n cannot meet the writes of scan n + 2, whose writers have passed the barrier of n + 1.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total, uint32_t& calls) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const tot = s_warp + 32 * (calls & 1u);
    calls++;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) tot[warp] = inc;
    __syncthreads();
    const uint32_t wv = tot[lane];
    uint32_t winc = wv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += t;
    }
    *total = __shfl_sync(0xffffffffu, winc, 31);
    return __shfl_sync(0xffffffffu, winc - wv, warp) + inc - v;
}

// ---------------------------------------------------------------------------
// K2: level-1 partition -> leaves -> sorted runs
// ---------------------------------------------------------------------------
// compare-and-swap on a leaf table slot (CUDA has the 64-bit overload for unsigned long long only)
__device__ __forceinline__ uint32_t leaf_cas(uint32_t* p, uint32_t cmp, uint32_t val) { return atomicCAS(p, cmp, val); }
__device__ __forceinline__ uint64_t leaf_cas(uint64_t* p, uint64_t cmp, uint64_t val) {
    return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)val);
}
__device__ __forceinline__ uint32_t leaf_hash(uint32_t key) { return key * 0x9E3779B1u; }
__device__ __forceinline__ uint32_t leaf_hash(uint64_t key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32); }

// Shared-memory geometry of the dedupe-first leaf pass (phase B of sp_leaf_kernel): a leaf may hold any number of
// records its scratch region takes, but at most D_MAX DISTINCT codes; the table has >= 1.5 slots per distinct code.
template <typename Shape, typename R2T>
struct LeafTable {
    static constexpr int D_MAX = Shape::LEAF_CAP * 4 / (int)sizeof(R2T) / 2;
    static constexpr int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
    static constexpr int SLOTS = pow2ceil(D_MAX * 3 / 2);
    static constexpr size_t BYTES = 1024 * 4 + (size_t)SLOTS * (sizeof(R2T) + 4) + (size_t)D_MAX * (sizeof(R2T) + 4);
};

template <typename Shape, typename R1T, typename R2T>
__global__ void __launch_bounds__(SP_THREADS, 1)
sp_leaf_kernel(int cb /* code bits of this round = 2k - rbits */, uint64_t prefix /* round << cb */, const R1T* __restrict__ slabs1, const uint32_t* __restrict__ counts1, uint32_t cap1,
               uint32_t grid1, uint32_t nsrc, uint32_t nparts, uint32_t part_first, R2T* __restrict__ scratch2,
               uint32_t cap2, uint64_t* __restrict__ tmp_keys,
               uint32_t* __restrict__ tmp_counts, uint64_t out_cap, unsigned long long* __restrict__ leaf_base,
               uint32_t* __restrict__ leaf_n, SpCtl* ctl) {
    KC_DYN_SMEM(uint32_t, smem);
    __shared__ uint32_t s_part, s_np;
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_warp[64];
    using St = Stager<R2T, Shape::P2>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r1bits = cb - Shape::KB1;
    const int r2bits = r1bits - Shape::KB2;  // >= 1 (host checks)
    const int dbits = r2bits < 10 ? r2bits : 10;
    const int lowbits = r2bits - dbits;
    const uint32_t nb = 1u << dbits;  // sub-buckets of a leaf, one per thread
    const uint64_t r2mask = (1ull << r2bits) - 1ull;
    R2T* const my_scratch = scratch2 + (uint64_t)blockIdx.x * Shape::P2 * cap2;
    // shared memory: the staging block (its cur[] holds the leaf record counts after Stager::finish)
    // phase B lies over the staging BINS (free once Stager::finish has run; state[] and cur[] stay):
    // hist/cursor[1024] | table keys[SLOTS] | table counts[SLOTS] | sorted keys[D_MAX] | sorted counts[D_MAX]
    using LT = LeafTable<Shape, R2T>;
    uint32_t* const s_hist = smem + 3 * Shape::P2;  // 1024 words
    R2T* const t_keys = reinterpret_cast<R2T*>(s_hist + 1024);
    uint32_t* const t_cnt = reinterpret_cast<uint32_t*>(t_keys + LT::SLOTS);
    R2T* const o_keys = reinterpret_cast<R2T*>(t_cnt + LT::SLOTS);
    uint32_t* const o_cnt = reinterpret_cast<uint32_t*>(o_keys + LT::D_MAX);
    __shared__ uint32_t s_special, s_leaf_fail, s_ok;

    uint32_t scans = 0;  // block scans made so far (block_excl_scan alternates its two total arrays)
    for (;;) {
        __syncthreads();  // previous partition fully done (s_part, the staging area and what phase B laid over it)
        if (tid == 0) s_part = atomicAdd(&ctl->work, 1u);
        St st;
        st.init(smem, my_scratch, (uint64_t)cap2, cap2, &ctl->failed);
        // records of this partition (sum over the pass-1 CTAs' regions)
        __syncthreads();
        // Local partition q of `nparts`; its records lie in nsrc * grid1 regions: source rank s sent
        // the block [s][q][cta][cap1] (single GPU: nsrc = 1, nparts = P1, part_first = 0).
        const uint32_t q1 = s_part;
        if (q1 >= nparts) break;
        const uint32_t p1 = part_first + q1;  // global partition = top KB1 bits of the codes in it
        const uint32_t nregions = nsrc * grid1;
        auto region_index = [&](uint32_t reg) {  // reg = s * grid1 + cta
            const uint32_t sr = reg / grid1, cta = reg - sr * grid1;
            return ((uint64_t)sr * nparts + q1) * grid1 + cta;
        };
        {
            uint32_t mine = 0;
            for (uint32_t r = tid; r < nregions; r += SP_THREADS) mine += counts1[region_index(r)];
            uint32_t tot;
            block_excl_scan(mine, s_warp, &tot, scans);
            if (tid == 0) s_np = tot;
        }
        __syncthreads();
        if (s_np == 0) {  // empty partition: its leaves keep leaf_n = 0 (zeroed by the host)
            continue;
        }

        // ---- phase A: partition by the next KB2 bits ------------------------------------
        // (the P2 mask only matters after an overflow upstream left garbage in a region: the
        // result is discarded then, but the kernel must stay in bounds)
        auto stage1 = [&](uint64_t r1, uint32_t pref) {
            st.stage((uint32_t)(r1 >> r2bits) & (Shape::P2 - 1), (R2T)(r1 & r2mask), pref);
        };
        // The staging code is long (it contains the bin flush), so it must not be inlined once
        // per record of a 128-bit load: the records of two loads sit in eight registers and ONE
        // copy of the staging code runs in a rolled loop that picks them with selects.
        auto pick = [](const uint32_t (&w)[8], int e) {
            uint32_t r = w[0];
#pragma unroll
            for (int q = 1; q < 8; q++) r = (e == q) ? w[q] : r;
            return r;
        };
        constexpr uint32_t RPV = 16 / sizeof(R1T);  // records per 128-bit load
        for (uint32_t reg = warp; reg < nregions; reg += SP_THREADS / 32) {
            // the result is void already: do not grind on (a short run list is not void: the host wants the full count of runs)
            // (decided by the whole warp: the loop below holds __syncwarp)
            if (__any_sync(0xffffffffu, (*(volatile uint32_t*)&ctl->failed & ~(uint32_t)SP_FAIL_RUNLIST) != 0)) break;
            const uint64_t ri = region_index(reg);
            const uint32_t n = counts1[ri];
            const R1T* src = slabs1 + ri * cap1;  // 64-byte aligned
            const uint4* src4 = reinterpret_cast<const uint4*>(src);
            const uint32_t nv = n / RPV;
            for (uint32_t i0 = 0; i0 < nv; i0 += 64) {  // two 128-bit loads in flight per lane; warp-uniform trip count
                const uint32_t i = i0 + (uint32_t)lane;
                const bool one = i < nv, two = i + 32 < nv;
                const uint4 v0 = one ? kc_ldg_stream(src4 + i) : make_uint4(0, 0, 0, 0);
                const uint4 v1 = two ? kc_ldg_stream(src4 + i + 32) : make_uint4(0, 0, 0, 0);
                const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                const int words = two ? 8 : (one ? 4 : 0);
#pragma unroll 1
                for (int e = 0; e < words; e += (int)sizeof(R1T) / 4) {
                    uint64_t r1 = pick(w, e);
                    if (sizeof(R1T) == 8) r1 |= (uint64_t)pick(w, e + 1) << 32;
                    stage1(r1, (uint32_t)warp + (uint32_t)e);
                }
                __syncwarp();  // the staging paths of the lanes end here (see kc_warp_scan): the next loads are issued by the whole warp
            }
            for (uint32_t t = nv * RPV + lane; t < n; t += 32) stage1((uint64_t)src[t], (uint32_t)warp);
        }
        __syncthreads();
        st.finish(smem, nullptr, 0, 0);
        __syncthreads();  // cur[p2] = records of leaf p2; scratch writes of this CTA are visible to it

        // ---- phase B: count the DISTINCT codes of a leaf first, sort those ---------------------------------
        // A leaf of deep-coverage reads holds ~coverage copies of every code in it.  Round 2's first form sorted the
        // RECORDS (counting sort on 10 bits, then a per-thread pass that collapsed each sub-bucket's duplicates): it paid
        // for every copy and every leaf waited for the thread with the largest sub-bucket — 93 K cycles per 12.4 K-record
        // leaf at config 4, 0.6 instructions issued per clock, leaf kernel 351 ms; a warp-cooperative collapse
        // (match.any) was slower still (395 ms).  Here every record is one probe of a shared-memory hash table (read
        // the slot; equal: count it; empty: claim it with a CAS) — the same work for every thread — and only the distinct
        // codes (1/7 of the records at 30 x) are counting-sorted on their top bits and finished by a per-thread insertion
        // sort of 1.7 entries on average.  No run-length pass: the entries are distinct, a leaf's run count is the
        // table's population and a thread's output offset its cursor.  Measured: 351 -> 151 ms (config 4), 190 -> 133 ms
        // (config 5, 50 M reads).  A leaf may now hold any number of records, but at most LeafTable::D_MAX distinct codes.
        const R2T EMPTY = ~(R2T)0;  // a real code only when r2bits fills R2T: counted in s_special instead
        for (uint32_t i = tid; i < (uint32_t)LT::SLOTS; i += SP_THREADS) {  // the bins of phase A lay here
            t_keys[i] = EMPTY;
            t_cnt[i] = 0;
        }
        s_hist[tid] = 0;
        if (tid == 0) s_special = 0, s_leaf_fail = 0;
        __syncthreads();
        for (uint32_t p2 = 0; p2 < (uint32_t)Shape::P2; p2++) {
            const uint32_t n = smem[2 * Shape::P2 + p2];
            if (n == 0) continue;  // CTA-uniform
            const R2T* leaf = my_scratch + (uint64_t)p2 * cap2;
            // table size for this leaf: two slots per record, at most SLOTS (slots outside stay clean)
            uint32_t slots = n <= 32 ? 64u : (2u << (31 - __clz((int)(2 * n - 1))));  // smallest power of two >= 2 n
            if (slots > (uint32_t)LT::SLOTS) slots = LT::SLOTS;
            const uint32_t smask = slots - 1u;
            const int hshift = 32 - (31 - __clz(slots));
            auto sub_of = [&](R2T key) { return (uint32_t)((uint64_t)key >> lowbits) & (nb - 1u); };
            auto insert = [&](R2T key) {
                if (key == EMPTY) {
                    KC_STAT(12);
                    if (atomicAdd(&s_special, 1u) == 0) atomicAdd(&s_hist[nb - 1u], 1u);
                    return;
                }
                uint32_t slot = (leaf_hash(key) >> hshift) & smask;
                for (uint32_t probes = 0;; probes++) {
                    R2T old = *(volatile R2T*)&t_keys[slot];
                    if (old == EMPTY) old = leaf_cas(&t_keys[slot], EMPTY, key);
                    if (old == EMPTY) {
                        atomicAdd(&s_hist[sub_of(key)], 1u);  // a new distinct code
                        break;
                    }
                    if (old == key) break;
                    slot = (slot + 1u) & smask;
                    if (probes >= 64u && (probes > smask || *(volatile uint32_t*)&s_leaf_fail)) {  // table full: the leaf fails
                        s_leaf_fail = 1;
                        return;
                    }
                }
                atomicAdd(&t_cnt[slot], 1u);
            };
            for (uint32_t i0 = 0; i0 < n; i0 += 4 * SP_THREADS) {  // four loads in flight per thread
                R2T r[4];
                bool have[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t i = i0 + (uint32_t)q * SP_THREADS + (uint32_t)tid;
                    have[q] = i < n;
                    r[q] = have[q] ? kc_ld_cg(leaf + i) : (R2T)0;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (have[q]) insert(r[q]);
                __syncwarp();  // probe loops differ per lane: the next four loads are issued by the whole warp
            }
            __syncthreads();
            const uint32_t cnt = (uint32_t)tid < nb ? s_hist[tid] : 0u;
            uint32_t nd;  // distinct codes of the leaf
            const uint32_t begin = block_excl_scan(cnt, s_warp, &nd, scans);
            if ((uint32_t)tid < nb) s_hist[tid] = begin;  // becomes the scatter cursor
            if (tid == 0) {
                uint32_t ok = 0;
                if (nd > (uint32_t)LT::D_MAX || s_leaf_fail) {
                    atomicOr(&ctl->failed, (uint32_t)SP_FAIL_LEAF);
                } else {
                    const unsigned long long base = atomicAdd(&ctl->out_cursor, (unsigned long long)nd);
                    s_base = base;
                    if (base + nd <= out_cap) {
                        leaf_base[(uint64_t)q1 * Shape::P2 + p2] = base;
                        leaf_n[(uint64_t)q1 * Shape::P2 + p2] = nd;
                        ok = 1;
                    } else {
                        atomicOr(&ctl->failed, (uint32_t)SP_FAIL_RUNLIST);  // temporary list full; leaf_n stays 0
                    }
                }
                s_ok = ok;
            }
            __syncthreads();
            const bool ok = s_ok != 0;
            // empty the table into the sorted area, sub-bucket by sub-bucket; the table is clean again afterwards
            for (uint32_t slot = tid; slot < slots; slot += SP_THREADS) {
                const R2T key = t_keys[slot];
                if (key != EMPTY) {
                    const uint32_t c = t_cnt[slot];
                    t_keys[slot] = EMPTY;
                    t_cnt[slot] = 0;
                    if (ok) {
                        const uint32_t pos = atomicAdd(&s_hist[sub_of(key)], 1u);
                        o_keys[pos] = key;
                        o_cnt[pos] = c;
                    }
                }
            }
            if (tid == 0 && s_special) {
                if (ok) {
                    const uint32_t pos = atomicAdd(&s_hist[nb - 1u], 1u);
                    o_keys[pos] = EMPTY;
                    o_cnt[pos] = s_special;
                }
                s_special = 0;
            }
            __syncthreads();
            if (tid == 0) s_leaf_fail = 0;
            if ((uint32_t)tid < nb) s_hist[tid] = 0;  // the next leaf's histogram
            if (ok && cnt > 1) {  // thread t owns sub-bucket t = o_[begin, begin + cnt): distinct codes, a handful
                for (uint32_t a = begin + 1; a < begin + cnt; a++) {
                    const R2T v = o_keys[a];
                    const uint32_t vc = o_cnt[a];
                    uint32_t b = a;
                    while (b > begin && o_keys[b - 1] > v) {
                        o_keys[b] = o_keys[b - 1];
                        o_cnt[b] = o_cnt[b - 1];
                        b--;
                    }
                    o_keys[b] = v;
                    o_cnt[b] = vc;
                }
            }
            __syncthreads();
            if (ok) {  // the leaf's runs, in code order, to the temporary list: coalesced
                const unsigned long long base = s_base;
                const uint64_t hi = prefix | ((uint64_t)p1 << r1bits) | ((uint64_t)p2 << r2bits);
                for (uint32_t i = tid; i < nd; i += SP_THREADS) {
                    tmp_keys[base + i] = hi | (uint64_t)o_keys[i];
                    tmp_counts[base + i] = o_cnt[i];
                }
            }
            // the next leaf writes o_[] and s_base only behind two more barriers
        }
    }
}

// ---------------------------------------------------------------------------
// K3: exclusive scan of leaf_n (leaves in code order), three small launches: per-block sums, scan of the block sums
// by one CTA, per-block scan with its offset.  (One CTA walking all 2^20 leaves took 1.8 ms: profiles/r02_ncu_sp_*.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SP_THREADS, 1)
sp_scan_sums_kernel(const uint32_t* __restrict__ leaf_n, uint64_t nleaves, unsigned long long* __restrict__ block_sums) {
    __shared__ uint32_t s_warp[64];
    const uint64_t i = (uint64_t)blockIdx.x * SP_THREADS + threadIdx.x;
    uint32_t tot, scans = 0;
    block_excl_scan(i < nleaves ? leaf_n[i] : 0u, s_warp, &tot, scans);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SP_THREADS, 1)
sp_scan_blocks_kernel(unsigned long long* __restrict__ block_sums, uint32_t nblocks, SpCtl* ctl) {
    // nblocks <= a few thousand: thread t owns a run of blocks, thread 0 combines the 1024 run totals
    __shared__ unsigned long long s_sum[SP_THREADS];
    const int tid = threadIdx.x;
    const uint32_t per = (nblocks + SP_THREADS - 1) / SP_THREADS;
    const uint32_t b = min((uint32_t)tid * per, nblocks), e = min(b + per, nblocks);
    unsigned long long mine = 0;
    for (uint32_t i = b; i < e; i++) mine += block_sums[i];
    s_sum[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < SP_THREADS; t++) {
            const unsigned long long v = s_sum[t];
            s_sum[t] = run;
            run += v;
        }
        ctl->total = run;
    }
    __syncthreads();
    unsigned long long run = s_sum[tid];
    for (uint32_t i = b; i < e; i++) {
        const unsigned long long v = block_sums[i];
        block_sums[i] = run;  // exclusive prefix of the block
        run += v;
    }
}

__global__ void __launch_bounds__(SP_THREADS, 1)
sp_scan_kernel(const uint32_t* __restrict__ leaf_n, uint64_t nleaves, const unsigned long long* __restrict__ block_sums,
               unsigned long long* __restrict__ leaf_off) {
    __shared__ uint32_t s_warp[64];
    const uint64_t i = (uint64_t)blockIdx.x * SP_THREADS + threadIdx.x;
    uint32_t tot, scans = 0;
    const uint32_t ex = block_excl_scan(i < nleaves ? leaf_n[i] : 0u, s_warp, &tot, scans);
    if (i < nleaves) leaf_off[i] = block_sums[blockIdx.x] + ex;
}

// ---------------------------------------------------------------------------
// K4: leaf runs -> final arrays (one warp per leaf at a time)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sp_gather_kernel(const uint32_t* __restrict__ leaf_n, const unsigned long long* __restrict__ leaf_base,
                 const unsigned long long* __restrict__ leaf_off, uint64_t nleaves, const uint64_t* __restrict__ tmp_keys,
                 const uint32_t* __restrict__ tmp_counts, uint64_t* __restrict__ keys, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const uint64_t nw = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t leaf = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); leaf < nleaves; leaf += nw) {
        const uint32_t n = leaf_n[leaf];
        if (n == 0) continue;
        const uint64_t src = leaf_base[leaf], dst = leaf_off[leaf];
        for (uint32_t i = lane; i < n; i += 32) {
            keys[dst + i] = tmp_keys[src + i];
            counts[dst + i] = tmp_counts[src + i];
        }
    }
}

struct DevMem {  // pool memory (kc_pool_alloc)
    void* p = nullptr;
    ~DevMem() { kc_pool_free(p); }
    cudaError_t alloc(size_t n) { return kc_pool_alloc(&p, n); }
    void* release() {
        void* q = p;
        p = nullptr;
        return q;
    }
};

// Records per private region for `mean` expected records: mean + 1/slack_div + sigmas * sqrt(mean)
// + 4 chunks, a whole number of chunks.  Level 1 regions (one per pass-1 CTA and partition)
// see near-Poisson counts; a LEAF holds whole families of repeated k-mers (coverage x copies
// of every genomic k-mer that falls into it), so its count is compound-Poisson with a far
// larger variance: level 2 gets twice the relative slack.
uint64_t region_records(uint64_t mean, int chunk, int slack_div, int sigmas) {
    uint64_t s = 1;
    while (s * s < mean) s++;
    uint64_t c = mean + mean / slack_div + sigmas * s + 4 * (uint64_t)chunk;
    return (c + chunk - 1) / chunk * chunk;
}

// The three stages of the host side.  A single GPU runs them back to back (kc_sparse_radix);
// the multi-GPU path runs `scatter` on every rank, exchanges the slab blocks of each rank's
// partition range with ONE all-to-all (kmerb200/distributed.py), and runs `count` on what it
// received — every rank then owns a disjoint, sorted range of codes and nothing is merged.
template <typename Shape, typename R1T, int HALO>
int run_scatter(kc_ctx* ctx, const char* d_data, uint64_t nbytes, const kc_radix_plan* plan, uint32_t round, void* d_slabs,
                uint32_t* d_counts, SpCtl* ctl) {
    using St1 = Stager<R1T, Shape::P1>;
    cudaStream_t st = ctx->stream;
    const int k = plan->k;
    const uint64_t nwin = nbytes >= (uint64_t)k ? nbytes - k + 1 : 0;
    KC_CUDA(ctx, cudaMemsetAsync(ctl, 0, sizeof(SpCtl), st));
    if (nwin == 0) {  // nothing to scan: every region is empty
        KC_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)plan->counts_bytes, st));
        return KC_OK;
    }
    const ScanGeom g = kc_make_geom(d_data, nbytes, 0, nwin, k);
    const size_t smem1 = St1::SMEM_BYTES;
    auto kern = sp_scatter_kernel<Shape, R1T, HALO>;
    KC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    KC_LAUNCH(kern, (int)plan->grid, SP_THREADS, smem1, st, g, (R1T*)d_slabs, d_counts, (uint32_t)plan->region_records, ctl,
              (int)plan->round_bits, round);
    KC_LAUNCH_CHECK(ctx, "sp_scatter_kernel");
    return KC_OK;
}

// Rounds on one GPU append their results (ascending code ranges) to ONE pair of arrays sized from the first round,
// instead of leaving a piece each to be concatenated: no second copy of the result in memory, no copy pass, and the
// same allocations in every call (the pool hands the same blocks out again).
struct SpAppend {
    uint64_t* keys = nullptr;
    uint32_t* counts = nullptr;
    uint64_t capacity = 0, used = 0;
    uint32_t rounds = 0;    // > 1: the first round with k-mers allocates the arrays, sized from its own count
    bool closed = false;    // a round did not fit: the later ones must stay behind it, as pieces
    bool appended = false;  // set by run_count: the round's result went behind `used` (its kc_sparse has no arrays)
};

template <typename Shape, typename R1T, typename R2T>
int run_count(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round, const void* d_slabs, const uint32_t* d_counts, uint32_t nsrc,
              uint32_t part_first, uint32_t nparts, char* work, size_t work_bytes, SpAppend* into, kc_sparse** out, int* failed) {
    using St2 = Stager<R2T, Shape::P2>;
    cudaStream_t st = ctx->stream;
    const int cb = 2 * plan->k - (int)plan->round_bits;
    const int grid2 = ctx->sm_count < (int)nparts ? ctx->sm_count : (int)nparts;
    const uint64_t part_mean = (uint64_t)nsrc * (plan->max_windows >> plan->round_bits) / Shape::P1 + 1;
    const uint64_t cap2 = region_records((part_mean + part_mean / 8) / Shape::P2 + 1, St2::CAP, 4, 16);
    if ((uint64_t)Shape::P2 * cap2 >= (1ull << 32)) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "sparse radix: leaf region too large");
    const uint64_t nleaves = (uint64_t)nparts * Shape::P2;
    auto pad = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const size_t b_ctl = pad(sizeof(SpCtl));
    const size_t b_leaf_n = pad(nleaves * 4), b_leaf_base = pad(nleaves * 8), b_leaf_off = pad(nleaves * 8);
    const size_t b_scratch2 = pad((size_t)grid2 * Shape::P2 * cap2 * sizeof(R2T));
    const size_t b_sums = pad(((nleaves + SP_THREADS - 1) / SP_THREADS) * 8);
    if (b_ctl + b_leaf_n + b_leaf_base + b_leaf_off + b_scratch2 + b_sums > work_bytes)
        return kc_set_error(ctx, KC_ERR_INVALID, "sparse radix: internal work area too small");
    SpCtl* ctl = (SpCtl*)work;
    uint32_t* leaf_n = (uint32_t*)(work + b_ctl);
    unsigned long long* leaf_base = (unsigned long long*)((char*)leaf_n + b_leaf_n);
    unsigned long long* leaf_off = (unsigned long long*)((char*)leaf_base + b_leaf_base);
    R2T* scratch2 = (R2T*)((char*)leaf_off + b_leaf_off);

    // Temporary run list.  Its size must not depend on how much memory happens to be free (a different size every round
    // makes the stream-ordered pool drop and re-create its blocks: 1.06 s for one allocation in round 2's config-5 runs),
    // and one entry per window would be ~7 x what deep-coverage reads need: a quarter of this rank's share of the
    // round's window positions, at most 45 % of what is free.  The leaf kernel keeps counting past the end of a list
    // that turns out too short (SpCtl::out_cursor), so the second try below has the exact size.
    size_t free_b = 0, total_b = 0;
    KC_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    free_b += kc_pool_idle_bytes(ctx->device);  // what earlier calls returned to the pool is ours to take again
    // (45 %: the final arrays need the same again — unless the rounds append to arrays that exist already)
    const uint64_t fit_cap = (uint64_t)(free_b / 100 * ((into && into->keys) ? 90 : 45)) / 12;
    const uint64_t round_windows = plan->max_windows >> plan->round_bits;
    const uint64_t most = (uint64_t)nsrc * (round_windows + round_windows / 8) / (plan->partitions / nparts) + (1u << 20);  // this rank's share of the round
    uint64_t out_cap = most / 4 + (1u << 20);
    if (out_cap > fit_cap) out_cap = fit_cap;
    static const char* cap_env = getenv("KC_SPARSE_RADIX_RUNLIST");  // tests: a first list this short
    if (cap_env && (uint64_t)atoll(cap_env) < out_cap) out_cap = (uint64_t)atoll(cap_env);
    if (out_cap < 1024) return kc_set_error(ctx, KC_ERR_NOMEM, "sparse radix: no device memory left for the run list");
    kc_trace(ctx, "count: enter", true);
    DevMem tkeys, tcounts;
    SpCtl h;
    for (int attempt = 0;; attempt++) {
        if (tkeys.alloc(out_cap * 8) || tcounts.alloc(out_cap * 4)) {
            cudaGetLastError();
            return kc_set_error(ctx, KC_ERR_NOMEM, "sparse radix: out of device memory for the run list");
        }
        kc_trace(ctx, "count: run-list alloc");
        KC_CUDA(ctx, cudaMemsetAsync(ctl, 0, sizeof(SpCtl), st));
        KC_CUDA(ctx, cudaMemsetAsync(leaf_n, 0, nleaves * 4, st));
        using LT = LeafTable<Shape, R2T>;
        const size_t bins = (size_t)Shape::P2 * 2 * 64;  // phase B overlays the staging bins
        const size_t smem2 = St2::SMEM_BYTES - bins + (bins > LT::BYTES ? bins : LT::BYTES);
        {
            auto kern = sp_leaf_kernel<Shape, R1T, R2T>;
            KC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            KC_LAUNCH(kern, grid2, SP_THREADS, smem2, st, cb, (uint64_t)round << cb, (const R1T*)d_slabs, d_counts, (uint32_t)plan->region_records,
                      plan->grid, nsrc, nparts, part_first, scratch2, (uint32_t)cap2, (uint64_t*)tkeys.p, (uint32_t*)tcounts.p,
                      out_cap, leaf_base, leaf_n, ctl);
            KC_LAUNCH_CHECK(ctx, "sp_leaf_kernel");
        }
        KC_CUDA(ctx, cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        KC_CUDA(ctx, cudaStreamSynchronize(st));
        if (h.failed == (uint32_t)SP_FAIL_RUNLIST && attempt == 0 && h.out_cursor <= fit_cap) {
            // only the list was too short and the slabs are untouched: count again with the size the kernel reported
            KC_STAT(13);
            kc_pool_free(tkeys.release());
            kc_pool_free(tcounts.release());
            out_cap = h.out_cursor;
            continue;
        }
        break;
    }
    if (!h.failed) {
        const uint32_t nblocks = (uint32_t)((nleaves + SP_THREADS - 1) / SP_THREADS);
        unsigned long long* block_sums = (unsigned long long*)((char*)scratch2 + b_scratch2);
        KC_LAUNCH(sp_scan_sums_kernel, nblocks, SP_THREADS, 0, st, leaf_n, nleaves, block_sums);
        KC_LAUNCH_CHECK(ctx, "sp_scan_sums_kernel");
        KC_LAUNCH(sp_scan_blocks_kernel, 1, SP_THREADS, 0, st, block_sums, nblocks, ctl);
        KC_LAUNCH_CHECK(ctx, "sp_scan_blocks_kernel");
        KC_LAUNCH(sp_scan_kernel, nblocks, SP_THREADS, 0, st, leaf_n, nleaves, block_sums, leaf_off);
        KC_LAUNCH_CHECK(ctx, "sp_scan_kernel");
        KC_CUDA(ctx, cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        KC_CUDA(ctx, cudaStreamSynchronize(st));
    }
    kc_trace(ctx, "count: leaf + scan kernels");
    if (h.failed) {
        *failed = (int)h.failed;
        return KC_OK;
    }
    kc_sparse* res = new kc_sparse();
    res->ctx = ctx;
    res->device = ctx->device;
    res->size = h.total;
    *out = res;
    if (into) into->appended = false;
    if (h.total == 0) return KC_OK;
    if (into && !into->keys && !into->closed && into->rounds > 1) {
        // the rounds split the codes by their top bits: size the whole result from this one (+ 12.5 %)
        uint64_t cap = h.total * into->rounds;
        cap += cap / 8 + 1024;
        void *dk = nullptr, *dc = nullptr;
        if (kc_pool_alloc(&dk, cap * 8) == cudaSuccess && kc_pool_alloc(&dc, cap * 4) == cudaSuccess) {
            into->keys = (uint64_t*)dk;
            into->counts = (uint32_t*)dc;
            into->capacity = cap;
            into->used = 0;
        } else {  // no room for that: pieces, concatenated at the end
            cudaGetLastError();
            kc_pool_free(dk);
            into->closed = true;
        }
    }
    if (into && into->keys && !into->closed && into->used + h.total <= into->capacity) {
        KC_LAUNCH(sp_gather_kernel, ctx->sm_count * 8, 256, 0, st, leaf_n, leaf_base, leaf_off, nleaves, (const uint64_t*)tkeys.p,
                  (const uint32_t*)tcounts.p, into->keys + into->used, into->counts + into->used);
        KC_LAUNCH_CHECK(ctx, "sp_gather_kernel");
        KC_CUDA(ctx, cudaStreamSynchronize(st));
        kc_trace(ctx, "count: gather kernel (appended)");
        into->used += h.total;
        into->appended = true;
        return KC_OK;
    }
    DevMem fk, fc;
    if (fk.alloc(h.total * 8) || fc.alloc(h.total * 4)) {
        cudaGetLastError();
        kc_sparse_free(res);
        *out = nullptr;
        return kc_set_error(ctx, KC_ERR_NOMEM, "sparse radix: out of device memory for %llu k-mers", h.total);
    }
    kc_trace(ctx, "count: result alloc");
    KC_LAUNCH(sp_gather_kernel, ctx->sm_count * 8, 256, 0, st, leaf_n, leaf_base, leaf_off, nleaves, (const uint64_t*)tkeys.p,
              (const uint32_t*)tcounts.p, (uint64_t*)fk.p, (uint32_t*)fc.p);
    KC_LAUNCH_CHECK(ctx, "sp_gather_kernel");
    KC_CUDA(ctx, cudaStreamSynchronize(st));
    kc_trace(ctx, "count: gather kernel");
    res->d_keys = (uint64_t*)fk.release();
    res->d_counts = (uint32_t*)fc.release();
    return KC_OK;
}

// bytes of the work area run_count needs (ctl + leaf directory + level-2 scratch)
template <typename Shape>
size_t count_work_bytes(const kc_ctx* ctx, const kc_radix_plan* plan, uint32_t nsrc, uint32_t nparts) {
    const int rec2 = (2 * plan->k - (int)plan->round_bits - Shape::KB1 - Shape::KB2) <= 32 ? 4 : 8;
    const int grid2 = ctx->sm_count < (int)nparts ? ctx->sm_count : (int)nparts;
    const uint64_t part_mean = (uint64_t)nsrc * (plan->max_windows >> plan->round_bits) / Shape::P1 + 1;
    const uint64_t cap2 = region_records((part_mean + part_mean / 8) / Shape::P2 + 1, 64 / rec2, 4, 16);
    const uint64_t nleaves = (uint64_t)nparts * Shape::P2;
    auto pad = [](size_t n) { return (n + 255) & ~(size_t)255; };
    return pad(sizeof(SpCtl)) + pad(nleaves * 4) + 2 * pad(nleaves * 8) + pad((size_t)grid2 * Shape::P2 * cap2 * rec2) +
           pad(((nleaves + SP_THREADS - 1) / SP_THREADS) * 8);
}

template <typename Shape>
int make_plan(kc_ctx* ctx, uint64_t max_windows, int k, uint32_t world, uint32_t shape_id, uint32_t min_round_bits, kc_radix_plan* plan) {
    if (2 * k - Shape::KB1 - Shape::KB2 < 1) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "sparse radix needs 2k > %d (k=%d)", Shape::KB1 + Shape::KB2, k);
    if (world < 1 || Shape::P1 % world) return kc_set_error(ctx, KC_ERR_INVALID, "sparse radix: world %u must divide %d partitions", world, Shape::P1);
    // ROUNDS.  A leaf (one of P1 x P2 code ranges) is sorted in shared memory and must fit it; the slabs of all
    // windows must fit the device.  Large inputs are therefore counted in 2^round_bits rounds: round r keeps the
    // windows whose top round_bits code bits are r, and the partition tree works on the remaining bits.  Config 5
    // (30 G window positions, 64-bit records): 192 GB of slabs -> 8 rounds on one GPU; 28.8 K positions per leaf against 14.7 K
    // (leaf rule below) -> at least 2 rounds on any number of GPUs.
    static const int rb_env = getenv("KC_SPARSE_RADIX_RBITS") ? atoi(getenv("KC_SPARSE_RADIX_RBITS")) : -1;  // tests
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = total_b = 0;
    }
    // memory this count can take again: the pool's idle blocks, the ctx's own scratch (the slabs of the last call
    // live there) and what the caller says its allocator has cached (the multi-GPU path keeps the slabs in torch tensors)
    free_b += kc_pool_idle_bytes(ctx->device) + ctx->scratch_bytes + ctx->scratch2_bytes + (size_t)ctx->caller_reusable_bytes;
    int rbits = (int)min_round_bits;
    if (2 * k - rbits - Shape::KB1 - Shape::KB2 < 1) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "sparse radix: %u round bits leave no code bits (k=%d)", min_round_bits, k);
    auto fill = [&](int rb) {  // the plan for rb round bits
        const int r1 = 2 * k - rb - Shape::KB1;
        memset(plan, 0, sizeof *plan);
        plan->k = k;
        plan->world = world;
        plan->partitions = Shape::P1;
        plan->parts_per_rank = Shape::P1 / world;
        plan->shape = shape_id;
        plan->round_bits = (uint32_t)rb;
        plan->rec_bytes = r1 <= 32 ? 4 : 8;
        plan->max_windows = max_windows;
        const uint64_t ngroups = (max_windows + k + 511 + 15) / 512 + 1;
        const uint64_t want1 = (ngroups + 31) / 32;
        plan->grid = (uint32_t)(want1 < 1 ? 1 : (want1 > (uint64_t)ctx->sm_count ? (uint64_t)ctx->sm_count : want1));
        plan->region_records = region_records((max_windows >> rb) / ((uint64_t)Shape::P1 * plan->grid) + 1, 64 / (int)plan->rec_bytes, 8, 8);
        plan->slab_bytes = (uint64_t)Shape::P1 * plan->grid * plan->region_records * plan->rec_bytes;
        plan->counts_bytes = (uint64_t)Shape::P1 * plan->grid * 4;
    };
    for (;; rbits++) {
        const int r1 = 2 * k - rbits - Shape::KB1, r2 = r1 - Shape::KB2;
        if (r2 < 1) {
            rbits = rbits ? rbits - 1 : 0;
            break;
        }
        if (rb_env >= 0 && min_round_bits == 0) {
            if (rbits >= rb_env) break;
            continue;
        }
        const uint64_t d_max = (uint64_t)Shape::LEAF_CAP / (r2 <= 32 ? 2 : 4);  // LeafTable::D_MAX for records of 4 or 8 bytes
        const uint64_t leaf_mean = ((uint64_t)world * max_windows >> rbits) / ((uint64_t)Shape::P1 * Shape::P2);
        // A leaf is limited by its DISTINCT codes (the leaf table), which the plan cannot know: it sees window
        // POSITIONS (reads: 130 of 151 hold a valid window) and assumes every code has ~2.3 copies or more: positions
        // per leaf <= 2.88 x D_MAX.  Config 4 at 30 x: 14.4 K positions, 12.4 K records, 1.7 K distinct against
        // 10 240, one round.  Shallower inputs fill a leaf table; that costs one scatter and a cut-short leaf pass and
        // is retried with one more round bit (kc_sparse_radix, count_sparse_radix_sharded) before the hash table takes over.
        const bool leaf_ok = leaf_mean * 25 <= d_max * 72;
        // Device memory: the slabs (+ the received copy when sharded), the work area with the leaf scratch, the run list
        // of one round and the result with the slack the appended rounds give it, the last two at an assumed one distinct
        // k-mer per six windows (configs 4 / 5: 1 in 7.3 / 5.5; run_count takes a quarter of the positions for the list
        // when that fits).  An input that is less redundant than that runs out of memory or run-list room in some round
        // and is counted again with the exact list size, or retried with more rounds.
        fill(rbits);
        const uint64_t rw = max_windows >> rbits;
        const uint64_t result = max_windows / 6 * 12, run_list = ((rw + rw / 8) / 6 + (2u << 20)) * 12;
        const uint64_t need = (plan->slab_bytes + plan->counts_bytes) * (world > 1 ? 2 : 1) + count_work_bytes<Shape>(ctx, plan, world, Shape::P1 / world) +
                              result + result / 8 + run_list;
        const bool mem_ok = free_b == 0 || need <= free_b / 20 * 19;
        if ((leaf_ok && mem_ok) || rbits >= 8) break;
    }
    fill(rbits);
    if (plan->region_records >= (1ull << 32)) return kc_set_error(ctx, KC_ERR_UNSUPPORTED, "sparse radix: region too large");
    return KC_OK;
}

template <typename Shape>
int scatter_dispatch(kc_ctx* ctx, const char* d_data, uint64_t nbytes, const kc_radix_plan* plan, uint32_t round, void* d_slabs,
                     uint32_t* d_counts, SpCtl* ctl) {
    if (plan->rec_bytes == 4)
        return plan->k <= 17 ? run_scatter<Shape, uint32_t, 1>(ctx, d_data, nbytes, plan, round, d_slabs, d_counts, ctl)
                             : run_scatter<Shape, uint32_t, 2>(ctx, d_data, nbytes, plan, round, d_slabs, d_counts, ctl);
    return run_scatter<Shape, uint64_t, 2>(ctx, d_data, nbytes, plan, round, d_slabs, d_counts, ctl);
}

template <typename Shape>
int count_dispatch(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round, const void* d_slabs, const uint32_t* d_counts, uint32_t nsrc,
                   uint32_t part_first, uint32_t nparts, char* work, size_t work_bytes, SpAppend* into, kc_sparse** out, int* failed) {
    const int r2 = 2 * plan->k - (int)plan->round_bits - Shape::KB1 - Shape::KB2;
    if (plan->rec_bytes == 4)
        return run_count<Shape, uint32_t, uint32_t>(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, work, work_bytes, into, out, failed);
    if (r2 <= 32)
        return run_count<Shape, uint64_t, uint32_t>(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, work, work_bytes, into, out, failed);
    return run_count<Shape, uint64_t, uint64_t>(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, work, work_bytes, into, out, failed);
}

using ShapeShipped = SpShape<10, 10, 20480>;
using ShapeSmall = SpShape<4, 4, 2048>;  // test shape: small inputs (and the CPU emulator) fill regions and leaves

uint32_t shape_from_env() {
    static const char* shape = getenv("KC_SPARSE_RADIX_SHAPE");
    return (shape && shape[0] == 's') ? 1u : 0u;
}

bool plan_ok(kc_ctx* ctx, const kc_radix_plan* plan) {
    return ctx && plan && plan->k >= 1 && plan->k <= KC_MAX_K && plan->shape <= 1 && plan->grid >= 1 && plan->region_records >= 1 &&
           plan->round_bits <= 8 &&
           plan->partitions == (plan->shape ? (uint32_t)ShapeSmall::P1 : (uint32_t)ShapeShipped::P1);
}

}  // namespace

extern "C" {

int kc_sparse_radix_plan_rounds(kc_ctx* ctx, uint64_t max_windows_per_rank, int k, uint32_t world, uint32_t min_round_bits,
                                kc_radix_plan* plan) {
    if (!ctx || !plan) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_K) return kc_set_error(ctx, KC_ERR_INVALID, "sparse k must be 1..%d, got %d", KC_MAX_K, k);
    if (min_round_bits > 8) return kc_set_error(ctx, KC_ERR_INVALID, "at most 8 round bits");
    DeviceGuard dg(ctx->device);
    const uint32_t sh = shape_from_env();
    return sh ? make_plan<ShapeSmall>(ctx, max_windows_per_rank, k, world, sh, min_round_bits, plan)
              : make_plan<ShapeShipped>(ctx, max_windows_per_rank, k, world, sh, min_round_bits, plan);
}
int kc_sparse_radix_plan(kc_ctx* ctx, uint64_t max_windows_per_rank, int k, uint32_t world, kc_radix_plan* plan) {
    return kc_sparse_radix_plan_rounds(ctx, max_windows_per_rank, k, world, 0, plan);
}

int kc_sparse_radix_scatter_round(kc_ctx* ctx, const char* d_data, uint64_t nbytes, const kc_radix_plan* plan, uint32_t round,
                                  void* d_slabs, uint32_t* d_counts) {
    if (!plan_ok(ctx, plan) || !d_slabs || !d_counts || (!d_data && nbytes) || round >= (1u << plan->round_bits))
        return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_radix_scatter: bad argument");
    const uint64_t nwin = nbytes >= (uint64_t)plan->k ? nbytes - plan->k + 1 : 0;
    if (nwin > plan->max_windows) return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_radix_scatter: %llu windows, plan made for %llu", (unsigned long long)nwin, (unsigned long long)plan->max_windows);
    DeviceGuard dg(ctx->device);
    int rc = kc_scratch2_reserve(ctx, 256);
    if (rc) return rc;
    SpCtl* ctl = (SpCtl*)ctx->scratch2;
    rc = plan->shape ? scatter_dispatch<ShapeSmall>(ctx, d_data, nbytes, plan, round, d_slabs, d_counts, ctl)
                     : scatter_dispatch<ShapeShipped>(ctx, d_data, nbytes, plan, round, d_slabs, d_counts, ctl);
    if (rc) return rc;
    SpCtl h;
    KC_CUDA(ctx, cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h.failed)
        return kc_set_error(ctx, KC_ERR_TABLE_FULL, "sparse radix scatter overflowed (skewed input):%s%s", (h.failed & 1) ? " partition region" : "",
                            (h.failed & 2) ? " staging bins" : "");
    return KC_OK;
}

static int count_round_into(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round, const void* d_slabs, const uint32_t* d_counts,
                            uint32_t nsrc, uint32_t part_first, uint32_t nparts, SpAppend* into, kc_sparse** out) {
    if (!out) return KC_ERR_INVALID;
    *out = nullptr;
    if (!plan_ok(ctx, plan) || !d_slabs || !d_counts || nsrc < 1 || nparts < 1 || (uint64_t)part_first + nparts > plan->partitions ||
        round >= (1u << plan->round_bits))
        return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_radix_count: bad argument");
    DeviceGuard dg(ctx->device);
    const size_t wb = plan->shape ? count_work_bytes<ShapeSmall>(ctx, plan, nsrc, nparts) : count_work_bytes<ShapeShipped>(ctx, plan, nsrc, nparts);
    int rc = kc_scratch2_reserve(ctx, wb);
    if (rc) return rc;
    int failed = 0;
    rc = plan->shape ? count_dispatch<ShapeSmall>(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, (char*)ctx->scratch2, wb, into, out, &failed)
                     : count_dispatch<ShapeShipped>(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, (char*)ctx->scratch2, wb, into, out, &failed);
    if (rc) return rc;
    if (failed)
        return kc_set_error(ctx, KC_ERR_TABLE_FULL, "sparse radix count overflowed (skewed input):%s%s%s%s", (failed & 1) ? " leaf region" : "",
                            (failed & 2) ? " staging bins" : "", (failed & 4) ? " leaf buffer" : "", (failed & 8) ? " run list" : "");
    return KC_OK;
}

int kc_sparse_radix_count_round(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round, const void* d_slabs, const uint32_t* d_counts,
                                uint32_t nsrc, uint32_t part_first, uint32_t nparts, kc_sparse** out) {
    return count_round_into(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, nullptr, out);
}

// The rounds of a plan, one call per round in ascending order, into ONE result: *acc is NULL before the first round
// and holds everything counted so far after each call.  The first round with k-mers sizes the arrays for all rounds
// (its count x rounds + 12.5 %: the rounds split the codes by their top bits); a later round that does not fit makes
// them grow (one copy).  Nothing is concatenated at the end and the allocations of a call repeat in the next one.
int kc_sparse_radix_count_round_append(kc_ctx* ctx, const kc_radix_plan* plan, uint32_t round, const void* d_slabs,
                                       const uint32_t* d_counts, uint32_t nsrc, uint32_t part_first, uint32_t nparts,
                                       kc_sparse** acc) {
    if (!acc || !plan) return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_radix_count_round_append: bad argument");
    const uint32_t rounds = 1u << plan->round_bits;
    SpAppend a;
    a.rounds = rounds;
    if (*acc) {
        if ((*acc)->ctx != ctx) return kc_set_error(ctx, KC_ERR_INVALID, "kc_sparse_radix_count_round_append: result of another context");
        a.keys = (*acc)->d_keys;
        a.counts = (*acc)->d_counts;
        a.used = (*acc)->size;
        a.capacity = (*acc)->capacity ? (*acc)->capacity : (*acc)->size;
        a.closed = a.keys == nullptr;  // an empty result so far: the round's own arrays are adopted below
    }
    kc_sparse* piece = nullptr;
    int rc = count_round_into(ctx, plan, round, d_slabs, d_counts, nsrc, part_first, nparts, &a, &piece);
    if (rc) return rc;
    DeviceGuard dg(ctx->device);
    if (!*acc) {
        *acc = new kc_sparse();
        (*acc)->ctx = ctx;
        (*acc)->device = ctx->device;
    }
    kc_sparse* r = *acc;
    if (a.appended) {  // run_count wrote behind a.used (and allocated the arrays if this was the first round with k-mers)
        r->d_keys = a.keys;
        r->d_counts = a.counts;
        r->capacity = a.capacity;
        r->size = a.used;
    } else if (piece && piece->size) {
        if (!r->d_keys) {  // nothing so far: the piece's arrays become the result's
            r->d_keys = piece->d_keys;
            r->d_counts = piece->d_counts;
            r->size = r->capacity = piece->size;
            piece->d_keys = nullptr;
            piece->d_counts = nullptr;
        } else {  // did not fit behind the others: grow to what the remaining rounds are likely to need as well
            KC_STAT(14);
            uint64_t cap = r->size + piece->size + (uint64_t)(rounds - 1 - (round < rounds ? round : rounds - 1)) * (piece->size + piece->size / 8);
            void *dk = nullptr, *dc = nullptr;
            if (kc_pool_alloc(&dk, cap * 8) != cudaSuccess || kc_pool_alloc(&dc, cap * 4) != cudaSuccess) {
                cudaGetLastError();
                kc_pool_free(dk);
                kc_sparse_free(piece);
                return kc_set_error(ctx, KC_ERR_NOMEM, "kc_sparse_radix_count_round_append: out of device memory for %llu k-mers", (unsigned long long)cap);
            }
            cudaStream_t st = ctx->stream;
            KC_CUDA(ctx, cudaMemcpyAsync(dk, r->d_keys, r->size * 8, cudaMemcpyDeviceToDevice, st));
            KC_CUDA(ctx, cudaMemcpyAsync(dc, r->d_counts, r->size * 4, cudaMemcpyDeviceToDevice, st));
            KC_CUDA(ctx, cudaMemcpyAsync((uint64_t*)dk + r->size, piece->d_keys, piece->size * 8, cudaMemcpyDeviceToDevice, st));
            KC_CUDA(ctx, cudaMemcpyAsync((uint32_t*)dc + r->size, piece->d_counts, piece->size * 4, cudaMemcpyDeviceToDevice, st));
            KC_CUDA(ctx, cudaStreamSynchronize(st));
            kc_pool_free(r->d_keys);
            kc_pool_free(r->d_counts);
            r->d_keys = (uint64_t*)dk;
            r->d_counts = (uint32_t*)dc;
            r->size += piece->size;
            r->capacity = cap;
        }
    }
    kc_sparse_free(piece);
    return KC_OK;
}

// single-round forms (plans with round_bits = 0)
int kc_sparse_radix_scatter(kc_ctx* ctx, const char* d_data, uint64_t nbytes, const kc_radix_plan* plan, void* d_slabs,
                            uint32_t* d_counts) {
    if (plan && plan->round_bits) return kc_set_error(ctx, KC_ERR_INVALID, "this plan has %u rounds: use kc_sparse_radix_scatter_round", 1u << plan->round_bits);
    return kc_sparse_radix_scatter_round(ctx, d_data, nbytes, plan, 0, d_slabs, d_counts);
}
int kc_sparse_radix_count(kc_ctx* ctx, const kc_radix_plan* plan, const void* d_slabs, const uint32_t* d_counts, uint32_t nsrc,
                          uint32_t part_first, uint32_t nparts, kc_sparse** out) {
    if (plan && plan->round_bits) {
        if (out) *out = nullptr;
        return kc_set_error(ctx, KC_ERR_INVALID, "this plan has %u rounds: use kc_sparse_radix_count_round", 1u << plan->round_bits);
    }
    return kc_sparse_radix_count_round(ctx, plan, 0, d_slabs, d_counts, nsrc, part_first, nparts, out);
}

// results of consecutive rounds (or any ascending pieces) -> one result; the pieces stay valid
// (consume = true, internal: every piece is freed as soon as it has been copied, so that the peak is the result + one piece)
static int sparse_concat(kc_ctx* ctx, kc_sparse** parts, uint32_t nparts, kc_sparse** out, bool consume) {
    if (!ctx || !out || (!parts && nparts)) return KC_ERR_INVALID;
    *out = nullptr;
    DeviceGuard dg(ctx->device);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nparts; i++) total += parts[i] ? parts[i]->size : 0;
    kc_sparse* res = new kc_sparse();
    res->ctx = ctx;
    res->device = ctx->device;
    res->size = total;
    if (total) {
        void *dk = nullptr, *dc = nullptr;
        if (kc_pool_alloc(&dk, total * 8) != cudaSuccess || kc_pool_alloc(&dc, total * 4) != cudaSuccess) {
            cudaGetLastError();
            kc_pool_free(dk);
            delete res;
            return kc_set_error(ctx, KC_ERR_NOMEM, "kc_sparse_concat: out of device memory for %llu k-mers", (unsigned long long)total);
        }
        res->d_keys = (uint64_t*)dk;
        res->d_counts = (uint32_t*)dc;
        uint64_t at = 0;
        for (uint32_t i = 0; i < nparts; i++) {
            if (!parts[i] || !parts[i]->size) continue;
            KC_CUDA(ctx, cudaMemcpyAsync(res->d_keys + at, parts[i]->d_keys, parts[i]->size * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            KC_CUDA(ctx, cudaMemcpyAsync(res->d_counts + at, parts[i]->d_counts, parts[i]->size * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            at += parts[i]->size;
            if (consume) {
                KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                kc_sparse_free(parts[i]);
                parts[i] = nullptr;
            }
        }
        KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *out = res;
    return KC_OK;
}
int kc_sparse_concat(kc_ctx* ctx, kc_sparse* const* parts, uint32_t nparts, kc_sparse** out) {
    return sparse_concat(ctx, const_cast<kc_sparse**>(parts), nparts, out, false);
}

}  // extern "C"

// Single GPU: plan + scatter + count with the slabs in the ctx's scratch.  Called by
// kc_count_sparse (sparse.cu).  *failed != 0: nothing was produced, the caller recounts.
int kc_sparse_radix(kc_ctx* ctx, const char* d_data, uint64_t nbytes, int k, kc_sparse** out, int* failed) {
    *failed = 0;
    *out = nullptr;
    auto pad = [](size_t n) { return (n + 255) & ~(size_t)255; };
    uint32_t min_bits = 0;
    int retries = 0;
    for (;;) {
        kc_radix_plan plan;
        int rc = kc_sparse_radix_plan_rounds(ctx, nbytes - k + 1, k, 1, min_bits, &plan);
        if (rc == KC_ERR_UNSUPPORTED && min_bits) {  // no code bits left for another round
            *failed = 1;
            return KC_OK;
        }
        if (rc) return rc;
        rc = kc_scratch_reserve(ctx, pad((size_t)plan.slab_bytes) + pad((size_t)plan.counts_bytes));
        if (rc) return rc;
        void* slabs = ctx->scratch;
        uint32_t* counts = (uint32_t*)((char*)ctx->scratch + pad((size_t)plan.slab_bytes));
        kc_trace(ctx, "radix: plan + scratch", true);
        const uint32_t rounds = 1u << plan.round_bits;
        std::vector<kc_sparse*> parts(rounds + 1, nullptr);  // [0] = what the rounds appended, [1 + r] = round r's own piece
        SpAppend acc;
        acc.rounds = rounds;
        for (uint32_t r = 0; r < rounds && rc == KC_OK; r++) {
            rc = kc_sparse_radix_scatter_round(ctx, d_data, nbytes, &plan, r, slabs, counts);
            kc_trace(ctx, "radix: scatter");
            if (rc == KC_OK) rc = count_round_into(ctx, &plan, r, slabs, counts, 1, 0, plan.partitions, rounds > 1 ? &acc : nullptr, &parts[1 + r]);
            kc_trace(ctx, "radix: count (incl. frees)", true);
            if (rc != KC_OK || rounds == 1) continue;
            if (acc.appended || (parts[1 + r] && parts[1 + r]->size == 0)) {  // the round's k-mers are in acc already (or it has none)
                kc_sparse_free(parts[1 + r]);
                parts[1 + r] = nullptr;
            } else if (!acc.closed) {
                KC_STAT(14);
                acc.closed = true;  // a round that did not fit: the later ones must stay behind it, as pieces
            }
        }
        if (acc.keys) {  // the appended rounds are the first piece
            kc_sparse* head = new kc_sparse();
            head->ctx = ctx;
            head->device = ctx->device;
            head->size = acc.used;
            head->capacity = acc.capacity;
            head->d_keys = acc.keys;
            head->d_counts = acc.counts;
            parts[0] = head;
        }
        if (rc == KC_OK) {
            uint32_t npieces = 0, last = 0;
            for (uint32_t i = 0; i <= rounds; i++)
                if (parts[i]) npieces++, last = i;
            if (npieces == 1) {
                *out = parts[last];
                parts[last] = nullptr;
            } else if (npieces == 0) {
                *out = new kc_sparse();
                (*out)->ctx = ctx;
                (*out)->device = ctx->device;
            } else {
                rc = sparse_concat(ctx, parts.data(), rounds + 1, out, true);
                if (rc == KC_ERR_NOMEM) {  // the slabs have done their work: the pieces and their concatenation get the room
                    kc_scratch_release(ctx);
                    rc = sparse_concat(ctx, parts.data(), rounds + 1, out, true);
                }
            }
        }
        for (kc_sparse* p : parts) kc_sparse_free(p);
        if (rc != KC_ERR_TABLE_FULL && rc != KC_ERR_NOMEM) return rc;
        // An overflow (the error text says which).  Denser-than-planned leaves or regions get ONE more round bit at a time
        // while the plan has room; inputs that are skewed rather than dense (poly-A: one leaf takes everything) are
        // hopeless for any number of rounds and go to the caller's fallback after two retries.
        if (retries >= 2 || plan.round_bits >= 8) {
            if (rc == KC_ERR_NOMEM) return rc;
            *failed = 1;
            return KC_OK;
        }
        retries++;
        KC_STAT(11);
        min_bits = plan.round_bits + 1;
        kc_scratch_release(ctx);  // the next plan's slabs are smaller: give the memory back before they are allocated
    }
}

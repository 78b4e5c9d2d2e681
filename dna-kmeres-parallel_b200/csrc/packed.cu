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
#include <sched.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"

// hostpack.cpp (plain g++): ASCII -> store layout on the host cores, one range per call
extern "C" uint32_t kc_host_pack_range(const char* data, uint64_t n, uint8_t* packed, uint32_t* badmask, int force_scalar);

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

// ---------------------------------------------------------------------------------------------------
// Counting from HOST memory through the packed form.  kc_count_dense_host moves 1 byte per base over
// PCIe and is bound by it (3.1 GB: 57 ms, while the GPU counts them in 4 ms).  Here the host cores turn
// the ASCII into the store's layout first, the GPU turns it back at HBM speed:
//
//   packer threads --items--> ring of pinned slots --H2D (copy stream)--> d_packed, mask stage / d_mask
//                                  --event--> mask_expand_kernel, unpack_kernel --> d_ascii --> dense path
//
// * item = 2^20 bases, claimed from one atomic counter (in order, so the ring fills front to back);
//   slot = 16 items = 4 MiB packed + 2 MiB mask; the ring holds 8 slots (pinned, kept in the ctx).
// * the validity bitmap crosses the bus SPARSE: a packer notes which 4 KiB blocks (32 K bases) of its item's
//   bitmap hold a set bit; the thread that finishes a slot moves the dirty blocks to the front of the slot's
//   mask area, and the slot sends [header: dirty-block map][dirty blocks] in one copy to a stage area on the
//   device, where mask_expand_kernel rebuilds the bitmap (clean blocks are zeroed there).  A slot with more
//   than a quarter of its blocks dirty sends its bitmap as it is.  A genome with an N every ~300 K bases costs
//   0.25 + ~0.02 bytes per base on the bus instead of 0.375.
// * a packer may write slot s only when slot s - RING has left the host (`released`, published by the
//   calling thread after cudaEventQuery of that slot's copy); the calling thread issues the copies of slot
//   s once it is finished.  No thread ever blocks inside the CUDA runtime.
// * every COUNT_SLOTS slots (2^28 bases) the compute stream waits for the copies so far, rebuilds bitmap and
//   ASCII image of those bases and counts the windows that END inside it, exactly like kc_count_dense_host.
// The result is kc_count_dense's: unpacking restores every valid byte and turns every invalid one into
// 'N', which resets a window like the byte it replaces.
namespace {

constexpr int HP_ITEMS_PER_SLOT = 16;
constexpr int HP_RING = 8;
constexpr int HP_COUNT_SLOTS = 16;             // slots per expand + unpack + count call
constexpr int HP_HDR_WORDS = 64;               // slot header: [0,16) dirty-block map per item, [16] mode, [17] dirty blocks
constexpr uint32_t HP_MODE_SPARSE = 0, HP_MODE_FULL = 1;

// Shape of the pipeline.  item = bases per packer item; a bitmap block = 1/32 of an item's bitmap (at least one
// word).  KC_HOSTPACK_ITEM is a test aid: small items drive a small input through many slots, ring wrap-arounds
// and count calls (values <= 1024 bases are rounded up to a multiple of 32, larger ones to a multiple of 1024).
struct HpShape {
    uint64_t item;            // bases per item
    uint32_t block_words;     // bitmap words per block
    uint32_t bpi;             // blocks per item (<= 32)
    uint64_t slot;            // bases per slot
    uint32_t blocks_per_slot;
    uint32_t max_sparse;      // a slot with more dirty blocks goes as a full bitmap
    size_t off_hdr, off_mask, slot_bytes;   // ring slot layout: [packed][header][bitmap]
    size_t stage_stride;      // device stage per slot: header + max_sparse blocks
};
static const HpShape& hp_shape() {
    static const HpShape v = [] {
        const char* e = getenv("KC_HOSTPACK_ITEM");
        uint64_t x = e ? strtoull(e, nullptr, 0) : 0;
        if (!x) x = 1ull << 20;
        x = x <= 1024 ? (x + 31) / 32 * 32 : (x + 1023) / 1024 * 1024;
        HpShape h;
        h.item = x;
        const uint64_t item_words = x / 32;
        h.block_words = item_words <= 32 ? 1u : (uint32_t)(item_words / 32);
        h.bpi = (uint32_t)(item_words / h.block_words);
        h.slot = x * HP_ITEMS_PER_SLOT;
        h.blocks_per_slot = h.bpi * HP_ITEMS_PER_SLOT;
        h.max_sparse = h.blocks_per_slot / 4;
        h.off_hdr = (size_t)(h.slot / 4);
        h.off_mask = h.off_hdr + HP_HDR_WORDS * 4;
        h.slot_bytes = h.off_mask + (size_t)(h.slot / 8);
        h.stage_stride = (HP_HDR_WORDS + (size_t)h.max_sparse * h.block_words) * 4;
        return h;
    }();
    return v;
}
#define HP_ITEM (hp_shape().item)

// Rebuild the bitmap words of the slots [slot0, slot0 + gridDim.x) from their stage records: CTA (s, j) does
// item j of slot s.  A FULL slot's bitmap was copied to its place directly.
__global__ void __launch_bounds__(256)
mask_expand_kernel(const uint32_t* __restrict__ stage, uint64_t stage_stride_words, uint64_t slot0, uint32_t block_words, uint32_t bpi,
                   uint64_t total_words, uint32_t* __restrict__ d_mask) {
    const uint64_t slot = slot0 + blockIdx.x;
    const uint32_t* hdr = stage + slot * stage_stride_words;
    if (hdr[16] != HP_MODE_SPARSE) return;
    const uint32_t j = blockIdx.y;
    uint32_t before = 0;  // dirty blocks of the items in front of this one
    for (uint32_t q = 0; q < j; q++) before += __popc(hdr[q]);
    const uint32_t map = hdr[j];
    const uint32_t* blocks = hdr + HP_HDR_WORDS;
    const uint64_t item_words = (uint64_t)block_words * bpi;
    const uint64_t w0 = (slot * HP_ITEMS_PER_SLOT + j) * item_words;  // first bitmap word of the item
    for (uint64_t w = threadIdx.x; w < item_words; w += blockDim.x) {
        if (w0 + w >= total_words) break;
        const uint32_t blk = (uint32_t)w / block_words, off = (uint32_t)w % block_words;  // w < 2^15 words per item
        uint32_t v = 0;
        if ((map >> blk) & 1u) {
            const uint32_t rank = before + __popc(map & ((1u << blk) - 1u));
            v = blocks[(uint64_t)rank * block_words + off];
        }
        d_mask[w0 + w] = v;
    }
}

int host_threads(int asked) {
    if (asked <= 0) {  // KC_HOSTPACK_THREADS: measurement aid for the automatic choice
        static const int env = getenv("KC_HOSTPACK_THREADS") ? atoi(getenv("KC_HOSTPACK_THREADS")) : 0;
        asked = env;
    }
    if (asked > 0) return asked > 256 ? 256 : asked;
    int n = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    if (n < 1) n = (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    // a container may be allowed less CPU time than the cores it can be scheduled on (cgroup v2 cpu.max =
    // "<quota> <period>" or "max <period>"): threads beyond the quota only get throttled
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        long long quota = 0, period = 0;
        if (fscanf(f, "%lld %lld", &quota, &period) == 2 && quota > 0 && period > 0) {
            const int q = (int)((quota + period - 1) / period);
            if (q >= 1 && q < n) n = q;
        }
        fclose(f);
    }
    n -= 1;  // the calling thread drives the copies
    if (n < 1) n = 1;
    return n > 64 ? 64 : n;
}

struct HostPackPipe {
    const HpShape& S = hp_shape();
    const char* h_data;
    uint64_t n;
    uint64_t nitems, nslots;
    uint8_t* ring;
    std::atomic<uint64_t> next_item{0};
    std::atomic<uint64_t> released{0};  // slots [0, released) have left the host
    std::atomic<int> stop{0};
    std::atomic<int> done[HP_RING];     // items packed
    std::atomic<int> ready[HP_RING];    // 1 = header written, dirty blocks compacted: the slot may be copied

    uint64_t items_of(uint64_t slot) const {
        const uint64_t first = slot * HP_ITEMS_PER_SLOT;
        return nitems - first < (uint64_t)HP_ITEMS_PER_SLOT ? nitems - first : (uint64_t)HP_ITEMS_PER_SLOT;
    }
    // the thread that packed the last item of a slot: count the dirty blocks, move them to the front
    void finish_slot(uint8_t* base, uint64_t items) {
        uint32_t* hdr = reinterpret_cast<uint32_t*>(base + S.off_hdr);
        uint32_t* mask = reinterpret_cast<uint32_t*>(base + S.off_mask);
        for (uint64_t j = items; j < (uint64_t)HP_ITEMS_PER_SLOT; j++) hdr[j] = 0;  // the last slot: no such items (the ring slot may hold an older map)
        uint32_t dirty = 0;
        for (int j = 0; j < HP_ITEMS_PER_SLOT; j++) dirty += (uint32_t)__builtin_popcount(hdr[j]);
        hdr[17] = dirty;
        if (dirty > S.max_sparse) {
            hdr[16] = HP_MODE_FULL;
            return;
        }
        hdr[16] = HP_MODE_SPARSE;
        uint32_t outb = 0;
        for (uint32_t blk = 0; blk < S.blocks_per_slot && outb < dirty; blk++) {
            if (!((hdr[blk / S.bpi] >> (blk % S.bpi)) & 1u)) continue;
            if (outb != blk) memmove(mask + (size_t)outb * S.block_words, mask + (size_t)blk * S.block_words, (size_t)S.block_words * 4);
            outb++;
        }
    }
    void work() {
        for (;;) {
            const uint64_t it = next_item.fetch_add(1, std::memory_order_relaxed);
            if (it >= nitems) return;
            const uint64_t slot = it / HP_ITEMS_PER_SLOT;
            for (int spins = 0; slot >= released.load(std::memory_order_acquire) + HP_RING; spins++) {
                if (stop.load(std::memory_order_relaxed)) return;
                // the ring is full = the bus is the limit (the wanted state): a slot takes ~75 us to leave
                if (spins < 64) std::this_thread::yield();
                else std::this_thread::sleep_for(std::chrono::microseconds(40));
            }
            if (stop.load(std::memory_order_relaxed)) return;
            const int r = (int)(slot % HP_RING);
            uint8_t* base = ring + (size_t)r * S.slot_bytes;
            const uint64_t j = it % HP_ITEMS_PER_SLOT;
            const uint64_t b = it * S.item;
            const uint64_t len = n - b < S.item ? n - b : S.item;
            uint8_t* packed = base + j * (S.item / 4);
            uint32_t* mask = reinterpret_cast<uint32_t*>(base + S.off_mask) + j * (S.item / 32);
            const uint64_t block_bases = (uint64_t)S.block_words * 32;
            uint32_t map = 0;
            for (uint32_t blk = 0; (uint64_t)blk * block_bases < len; blk++) {
                const uint64_t o = (uint64_t)blk * block_bases;
                const uint64_t l = len - o < block_bases ? len - o : block_bases;
                if (kc_host_pack_range(h_data + b + o, l, packed + o / 4, mask + o / 32, 0)) map |= 1u << blk;
            }
            reinterpret_cast<uint32_t*>(base + S.off_hdr)[j] = map;
            // acq_rel: the finisher sees every item's bytes and map; its own are ordered before the flag
            if ((uint64_t)done[r].fetch_add(1, std::memory_order_acq_rel) + 1 == items_of(slot)) {
                finish_slot(base, items_of(slot));
                ready[r].store(1, std::memory_order_release);
            }
        }
    }
};

}  // namespace

extern "C" {

int kc_pack_2bit_host_body(const char* h_data, uint64_t nbytes, void* h_packed, uint32_t* h_badmask, int nthreads, int body) {
    if (nbytes == 0) return KC_OK;
    if (!h_data || !h_packed || !h_badmask) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_pack_2bit_host: null pointer");
    if (body < 0 || body > 2) return kc_set_error(nullptr, KC_ERR_INVALID, "kc_pack_2bit_host_body: body must be 0, 1 or 2");
    const uint64_t nitems = (nbytes + HP_ITEM - 1) / HP_ITEM;
    int T = host_threads(nthreads);
    if ((uint64_t)T > nitems) T = (int)nitems;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        for (;;) {
            const uint64_t it = next.fetch_add(1, std::memory_order_relaxed);
            if (it >= nitems) return;
            const uint64_t b = it * HP_ITEM;
            const uint64_t len = nbytes - b < HP_ITEM ? nbytes - b : HP_ITEM;
            kc_host_pack_range(h_data + b, len, (uint8_t*)h_packed + b / 4, h_badmask + b / 32, body);
        }
    };
    std::vector<std::thread> th;
    try {
        for (int t = 1; t < T; t++) th.emplace_back(work);
    } catch (...) {  // fewer threads than asked for: the ones that run (and this one) do all items
    }
    work();
    for (auto& t : th) t.join();
    return KC_OK;
}

int kc_host_pack_threads(int nthreads) { return host_threads(nthreads); }

int kc_pack_2bit_host(const char* h_data, uint64_t nbytes, void* h_packed, uint32_t* h_badmask, int nthreads) {
    return kc_pack_2bit_host_body(h_data, nbytes, h_packed, h_badmask, nthreads, 0);
}

}  // extern "C"

// h_table != NULL: the table ends in host memory (d_user NULL);  d_user != NULL: it is counted into the caller's
// device table (overwritten) and stays there
static int host_packed_count(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* h_table, uint32_t* d_user,
                             int nthreads) {
    if (!ctx) return KC_ERR_INVALID;
    if (k < 1 || k > KC_MAX_DENSE_K) return kc_set_error(ctx, KC_ERR_INVALID, "dense k must be 1..%d, got %d", KC_MAX_DENSE_K, k);
    if ((!h_table && !d_user) || (!h_data && nbytes)) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    DeviceGuard dg(ctx->device);
    const size_t table_bytes = sizeof(uint32_t) << (2 * k);
    // device image: table | ASCII (rounded up to 256 B) | packed | mask
    const size_t ascii_bytes = (size_t)((nbytes + 64 + 255) & ~(uint64_t)255);
    const size_t packed_bytes = (size_t)(((nbytes + 3) / 4 + 255) & ~(uint64_t)255);
    const size_t mask_bytes = (size_t)(((nbytes + 31) / 32 * 4 + 255) & ~(uint64_t)255);
    const HpShape& S = hp_shape();
    const uint64_t nitems_all = (nbytes + S.item - 1) / S.item;
    const uint64_t nslots_all = (nitems_all + HP_ITEMS_PER_SLOT - 1) / HP_ITEMS_PER_SLOT;
    const size_t stage_bytes = (size_t)nslots_all * S.stage_stride;
    int rc = kc_scratch2_reserve(ctx, table_bytes + ascii_bytes + packed_bytes + mask_bytes + stage_bytes + 256);
    if (rc) return rc;
    uint32_t* d_table = d_user ? d_user : (uint32_t*)ctx->scratch2;
    char* d_ascii = (char*)ctx->scratch2 + table_bytes;
    uint8_t* d_packed = (uint8_t*)d_ascii + ascii_bytes;
    uint32_t* d_mask = (uint32_t*)(d_packed + packed_bytes);
    uint32_t* d_stage = (uint32_t*)((uint8_t*)d_mask + mask_bytes);
    const uint64_t total_words = (nbytes + 31) / 32;
    KC_CUDA(ctx, cudaMemsetAsync(d_table, 0, table_bytes, ctx->stream));
    ctx->last_h2d_bytes = 0;
    if (nbytes >= (uint64_t)k) {
        const size_t ring_bytes = (size_t)HP_RING * S.slot_bytes;
        if (ctx->pinned_bytes < ring_bytes) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
            cudaError_t e = cudaHostAlloc(&ctx->pinned, ring_bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) {
                cudaGetLastError();
                ctx->pinned = nullptr;
                return kc_set_error(ctx, KC_ERR_NOMEM, "cudaHostAlloc(%zu): %s", ring_bytes, cudaGetErrorString(e));
            }
            ctx->pinned_bytes = ring_bytes;
        }
        HostPackPipe pipe;
        pipe.h_data = h_data;
        pipe.n = nbytes;
        pipe.nitems = nitems_all;
        pipe.nslots = nslots_all;
        pipe.ring = (uint8_t*)ctx->pinned;
        for (int i = 0; i < HP_RING; i++) {
            pipe.done[i].store(0, std::memory_order_relaxed);
            pipe.ready[i].store(0, std::memory_order_relaxed);
        }
        cudaEvent_t ev[HP_RING];
        int nev = 0;
        cudaError_t ce = cudaSuccess;
        for (; nev < HP_RING && ce == cudaSuccess; nev++) ce = cudaEventCreateWithFlags(&ev[nev], cudaEventDisableTiming);
        if (ce != cudaSuccess) {
            for (int i = 0; i + 1 < nev; i++) cudaEventDestroy(ev[i]);
            return kc_set_error(ctx, KC_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(ce));
        }
        int T = host_threads(nthreads);
        if ((uint64_t)T > pipe.nitems) T = (int)pipe.nitems;
        // Starting a thread costs ~20-30 us: 64 of them up front would keep this thread away from the ring for
        // 1-2 ms (10 % of a 3.1 Gbp call) while the first slots sit there finished.  A few start now, the rest one
        // per turn of the loop below, between looks at the ring.
        std::vector<std::thread> th;
        th.reserve((size_t)T);
        bool can_spawn = true;
        auto spawn = [&]() {
            try {
                th.emplace_back([&pipe]() { pipe.work(); });
            } catch (...) {  // fewer threads than wanted: the ones that run do all items
                can_spawn = false;
            }
        };
        for (int t = 0; t < T && t < 4 && can_spawn; t++) spawn();
        const uint64_t nwin = nbytes - k + 1;
        uint64_t issued = 0, released = 0, unpacked = 0, counted = 0, h2d = 0;
        rc = KC_OK;
        if (th.empty()) rc = kc_set_error(ctx, KC_ERR_NOMEM, "kc_count_dense_host_packed: no packer thread could be started");
        while (!rc && released < pipe.nslots) {
            bool progress = false;
            if (can_spawn && (int)th.size() < T && pipe.next_item.load(std::memory_order_relaxed) < pipe.nitems) {
                spawn();
                progress = true;
            }
            if (issued < pipe.nslots && issued < released + HP_RING && pipe.ready[issued % HP_RING].load(std::memory_order_acquire)) {
                const int r = (int)(issued % HP_RING);
                pipe.ready[r].store(0, std::memory_order_relaxed);  // both before `released` lets anyone at this slot again
                pipe.done[r].store(0, std::memory_order_relaxed);
                const uint64_t b = issued * S.slot;
                const uint64_t e = (b + S.slot < nbytes) ? b + S.slot : nbytes;
                const uint8_t* src = pipe.ring + (size_t)r * S.slot_bytes;
                const uint32_t* hdr = reinterpret_cast<const uint32_t*>(src + S.off_hdr);
                const bool sparse = hdr[16] == HP_MODE_SPARSE;
                // [header][dirty blocks] -> the slot's stage record (a FULL slot: the header alone)
                const size_t rec_bytes = (HP_HDR_WORDS + (sparse ? (size_t)hdr[17] * S.block_words : 0)) * 4;
                const size_t full_bytes = (size_t)((e - b + 31) / 32 * 4);
                ce = cudaMemcpyAsync(d_packed + b / 4, src, (size_t)((e - b + 3) / 4), cudaMemcpyHostToDevice, ctx->copy_stream);
                if (ce == cudaSuccess)
                    ce = cudaMemcpyAsync((uint8_t*)d_stage + issued * S.stage_stride, hdr, rec_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
                if (ce == cudaSuccess && !sparse)
                    ce = cudaMemcpyAsync(d_mask + b / 32, src + S.off_mask, full_bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
                h2d += (e - b + 3) / 4 + rec_bytes + (sparse ? 0 : full_bytes);
                if (ce == cudaSuccess) ce = cudaEventRecord(ev[r], ctx->copy_stream);
                issued++;
                if (ce == cudaSuccess && (issued % HP_COUNT_SLOTS == 0 || issued == pipe.nslots)) {
                    // bases [unpacked, e) are on their way: rebuild their bitmap and ASCII image behind the copies, count
                    ce = cudaStreamWaitEvent(ctx->stream, ev[r], 0);
                    if (ce == cudaSuccess) {
                        const uint64_t slot0 = unpacked / S.slot;
                        KC_LAUNCH(mask_expand_kernel, dim3((unsigned)(issued - slot0), HP_ITEMS_PER_SLOT), 256, 0, ctx->stream, d_stage,
                                  (uint64_t)(S.stage_stride / 4), slot0, S.block_words, S.bpi, total_words, d_mask);
                        ctx->launches++;
                        ce = cudaGetLastError();
                    }
                    if (ce == cudaSuccess) {
                        rc = unpack_range(ctx, d_packed, d_mask, unpacked, e - unpacked, d_ascii + unpacked, ctx->stream);
                        unpacked = e;
                        uint64_t upto = (e >= (uint64_t)k) ? e - k + 1 : 0;  // windows that end before byte e
                        if (upto > nwin) upto = nwin;
                        if (!rc && upto > counted) {
                            rc = kc_count_dense_range_async(ctx, d_ascii, e, counted, upto, k, d_table, KC_DENSE_AUTO, ctx->stream);
                            counted = upto;
                        }
                    }
                }
                if (ce != cudaSuccess) rc = kc_set_error(ctx, KC_ERR_CUDA, "host staging failed: %s", cudaGetErrorString(ce));
                progress = true;
            }
            if (!rc && released < issued) {
                ce = cudaEventQuery(ev[released % HP_RING]);
                if (ce == cudaSuccess) {
                    released++;
                    pipe.released.store(released, std::memory_order_release);
                    progress = true;
                } else if (ce == cudaErrorNotReady) {
                    cudaGetLastError();  // "not ready" is recorded as the thread's last error: the next launch check must not see it
                } else {
                    rc = kc_set_error(ctx, KC_ERR_CUDA, "host staging failed: %s", cudaGetErrorString(ce));
                }
            }
            if (!progress) std::this_thread::yield();
        }
        pipe.stop.store(1, std::memory_order_relaxed);
        pipe.next_item.store(pipe.nitems, std::memory_order_relaxed);
        for (auto& t : th) t.join();
        if (rc) cudaStreamSynchronize(ctx->copy_stream);  // the ring must not be reused under a copy in flight
        for (int i = 0; i < HP_RING; i++) cudaEventDestroy(ev[i]);
        if (rc) return rc;
        ctx->last_h2d_bytes = h2d;
    }
    if (h_table) KC_CUDA(ctx, cudaMemcpyAsync(h_table, d_table, table_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KC_OK;
}

// (A hybrid of the two host paths — the first 15-35 % of the input as plain bytes on a second ctx and host thread while
// the packers work on the rest — was built, verified and measured on the 16-core B200 host: 37.1 / 39.3 / 41.7 ms
// against 33.9 ms for the packed path alone.  The packers are bound by the host's memory system, which the DMA of the
// plain part loads further; removed again.  DESIGN.md section 7.)

extern "C" {

int kc_count_dense_host_packed(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* h_table, int nthreads) {
    if (ctx && !h_table) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    return host_packed_count(ctx, h_data, nbytes, k, h_table, nullptr, nthreads);
}

int kc_count_dense_host_packed_dev(kc_ctx* ctx, const char* h_data, uint64_t nbytes, int k, uint32_t* d_table, int nthreads) {
    if (ctx && !d_table) return kc_set_error(ctx, KC_ERR_INVALID, "null pointer");
    return host_packed_count(ctx, h_data, nbytes, k, nullptr, d_table, nthreads);
}

}  // extern "C"

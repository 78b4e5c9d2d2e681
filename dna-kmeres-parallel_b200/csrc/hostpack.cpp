// hostpack.cpp — HOST half of the packed host->device path (kc_count_dense_host_packed, packed.cu) and
// of the 2-bit store of "next" row f4: ASCII bases -> the byte layout the reference sketches in
// main.cu:78-86 / utils.h:65-92 (four bases per byte, the first base in the two most significant
// bits, A=00 C=01 G=10 T=11) + the validity bitmap of packed.cu (bit i%32 of word i/32 set iff byte i
// is not an upper-case ACGT; such bases pack as 00).  Format conversion only: nothing here counts
// k-mers — counting has no CPU path.
//
// Why it exists: counting from HOST memory is bound by PCIe (1 byte per base, ~55 GB/s), not by the
// GPU (770 Gbases/s).  Packing on the host cores first sends 0.375 bytes per base; the GPU unpacks at
// HBM speed.  Plain g++ (no CUDA): AVX-512BW or AVX2 body chosen at run time, scalar body otherwise and for tails.
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Lut {
    uint8_t v[256];
    Lut() {
        memset(v, 4, sizeof v);
        v['A'] = 0;
        v['C'] = 1;
        v['G'] = 2;
        v['T'] = 3;
    }
};
const Lut g_lut;

// bases [0, n) -> packed[0, (n+3)/4), mask[0, (n+31)/32); bits past n are set, bases past n pack as 00;
// returns the OR of the mask words written (0 = every base valid)
uint32_t pack_scalar(const uint8_t* s, uint64_t n, uint8_t* packed, uint32_t* mask) {
    uint32_t any = 0;
    const uint64_t full = n >> 5;
    for (uint64_t w = 0; w < full; w++) {
        const uint8_t* p = s + (w << 5);
        uint32_t bad = 0;
        for (int q = 0; q < 8; q++) {
            uint32_t byte = 0;
            for (int t = 0; t < 4; t++) {
                uint32_t c = g_lut.v[p[4 * q + t]];
                bad |= (c >> 2) << (4 * q + t);
                byte = (byte << 2) | (c & 3u & ~(0u - (c >> 2)));
            }
            packed[(w << 3) + q] = (uint8_t)byte;
        }
        mask[w] = bad;
        any |= bad;
    }
    const uint64_t rest = n & 31;
    if (rest) {
        const uint8_t* p = s + (full << 5);
        uint32_t bad = 0xFFFFFFFFu << rest;
        const uint64_t nb = (rest + 3) >> 2;
        for (uint64_t q = 0; q < nb; q++) {
            uint32_t byte = 0;
            for (uint64_t t = 0; t < 4; t++) {
                const uint64_t j = 4 * q + t;
                uint32_t c = j < rest ? g_lut.v[p[j]] : 0u;
                if (c > 3) {
                    bad |= 1u << j;
                    c = 0;
                }
                byte = (byte << 2) | c;
            }
            packed[(full << 3) + q] = (uint8_t)byte;
        }
        mask[full] = bad;
        any |= bad;
    }
    return any;
}

#if defined(__x86_64__)
// 32 bases per iteration.  Two byte shuffles indexed by the LOW NIBBLE of each input byte (A C G T = 0x41 0x43
// 0x47 0x54 have the distinct low nibbles 1 3 7 4): one returns the only letter that nibble may belong to — the
// byte is valid iff it equals it (unused entries hold 0xFF, which no byte with bit 7 clear equals; a byte with
// bit 7 set shuffles to 0, which it does not equal either) — the other its 2-bit code.  Two multiply-adds fold
// four codes into one byte, first base on top.
__attribute__((target("avx2"))) uint32_t pack_avx2(const uint8_t* s, uint64_t nwords, uint8_t* packed, uint32_t* mask) {
    const char F = (char)0xFF;
    const __m256i letter = _mm256_setr_epi8(F, 0x41, F, 0x43, 0x54, F, F, 0x47, F, F, F, F, F, F, F, F,
                                            F, 0x41, F, 0x43, 0x54, F, F, 0x47, F, F, F, F, F, F, F, F);
    const __m256i codes = _mm256_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0,
                                           0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i m1 = _mm256_set1_epi16(0x0104);      // even byte * 4 + odd byte
    const __m256i m2 = _mm256_set1_epi32(0x00010010);  // even half * 16 + odd half
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint32_t any = 0;
    for (uint64_t w = 0; w < nwords; w++) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + (w << 5)));
        const __m256i valid = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(letter, x), x);
        const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(codes, x), valid);
        const __m256i t = _mm256_maddubs_epi16(code, m1);
        const __m256i u = _mm256_madd_epi16(t, m2);
        const __m256i v = _mm256_shuffle_epi8(u, pick);
        const uint64_t lo = (uint32_t)_mm256_extract_epi32(v, 0), hi = (uint32_t)_mm256_extract_epi32(v, 4);
        const uint64_t out = lo | (hi << 32);
        memcpy(packed + (w << 3), &out, 8);
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(valid);
        mask[w] = bad;
        any |= bad;
    }
    return any;
}

// 64 bases per iteration, the same two nibble-indexed shuffles: the validity compare lands in a mask register
// (= the bitmap word pair, inverted), the code shuffle is zero-masked by it, and vpmovdb gathers the 16 result
// bytes in order.
__attribute__((target("avx512f,avx512bw"))) uint32_t pack_avx512(const uint8_t* s, uint64_t npairs, uint8_t* packed, uint32_t* mask) {
    const __m512i letter = _mm512_broadcast_i32x4(_mm_setr_epi8((char)0xFF, 0x41, (char)0xFF, 0x43, 0x54, (char)0xFF, (char)0xFF, 0x47,
                                                                (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF, (char)0xFF,
                                                                (char)0xFF, (char)0xFF));
    const __m512i codes = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i m1 = _mm512_set1_epi16(0x0104);
    const __m512i m2 = _mm512_set1_epi32(0x00010010);
    uint64_t any = 0;
    for (uint64_t w = 0; w < npairs; w++) {
        const __m512i x = _mm512_loadu_si512(s + (w << 6));
        const __mmask64 valid = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(letter, x), x);
        const __m512i code = _mm512_maskz_shuffle_epi8(valid, codes, x);
        const __m512i u = _mm512_madd_epi16(_mm512_maddubs_epi16(code, m1), m2);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(packed + (w << 4)), _mm512_cvtepi32_epi8(u));
        const uint64_t bad = ~(uint64_t)valid;
        memcpy(mask + (w << 1), &bad, 8);
        any |= bad;
    }
    return (uint32_t)(any | (any >> 32));
}
#endif

}  // namespace

extern "C" {

// 0 = scalar body, 1 = AVX2, 2 = AVX-512BW on this host
__attribute__((visibility("default"))) int kc_host_pack_simd(void) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) return 2;
    return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    return 0;
#endif
}

// One range, one thread: bases data[0, n) -> packed[0, (n+3)/4) and badmask[0, (n+31)/32).  A caller that
// splits a sequence cuts it at multiples of 32 bases (whole mask words, whole packed bytes).
// force_scalar: 0 = the best body this host has (AVX-512BW, AVX2, scalar), 1 = scalar, 2 = AVX2 (tests compare them).  Returns the OR of the mask words (0 = all valid).
__attribute__((visibility("hidden"))) uint32_t kc_host_pack_range(const char* data, uint64_t n, uint8_t* packed, uint32_t* badmask,
                                                                int force_scalar) {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(data);
    uint64_t done = 0;
    uint32_t any = 0;
#if defined(__x86_64__)
    // force_scalar: 0 = best body of this host, 1 = scalar, 2 = AVX2 even where AVX-512 exists (tests)
    if (force_scalar != 1 && force_scalar != 2 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f")) {
        const uint64_t npairs = n >> 6;
        any = pack_avx512(s, npairs, packed, badmask);
        done = npairs << 6;
    }
    if (force_scalar != 1 && __builtin_cpu_supports("avx2")) {
        const uint64_t nwords = (n - done) >> 5;
        any |= pack_avx2(s + done, nwords, packed + (done >> 2), badmask + (done >> 5));
        done += nwords << 5;
    }
#endif
    if (done < n) any |= pack_scalar(s + done, n - done, packed + (done >> 2), badmask + (done >> 5));
    return any;
}

}  // extern "C"

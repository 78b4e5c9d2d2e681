// cub_emu.h — host stand-ins for the few CUB device-wide primitives sparse.cu calls, so that
// the file builds against the test-only CPU emulator (simt_emu.h).  Same calling convention
// (a first call with d_temp_storage == nullptr returns the scratch size).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
struct Sum {};

struct DeviceRadixSort {
    template <class K>
    static uint64_t digit(K k, int b, int e) {
        const int w = e - b;
        const uint64_t m = w >= 64 ? ~0ull : ((1ull << w) - 1ull);
        return ((uint64_t)k >> b) & m;
    }
    template <class K, class V, class N>
    static cudaError_t SortPairs(void* tmp, size_t& tb, const K* kin, K* kout, const V* vin, V* vout, N n, int b = 0,
                                 int e = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!tmp) {
            tb = 16;
            return cudaSuccess;
        }
        std::vector<size_t> idx((size_t)n);
        std::iota(idx.begin(), idx.end(), (size_t)0);
        std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return digit(kin[x], b, e) < digit(kin[y], b, e); });
        for (size_t i = 0; i < (size_t)n; i++) {
            kout[i] = kin[idx[i]];
            vout[i] = vin[idx[i]];
        }
        return cudaSuccess;
    }
    template <class K, class N>
    static cudaError_t SortKeys(void* tmp, size_t& tb, const K* kin, K* kout, N n, int b = 0, int e = sizeof(K) * 8,
                                cudaStream_t = nullptr) {
        if (!tmp) {
            tb = 16;
            return cudaSuccess;
        }
        std::vector<K> v(kin, kin + (size_t)n);
        std::stable_sort(v.begin(), v.end(), [&](K x, K y) { return digit(x, b, e) < digit(y, b, e); });
        std::copy(v.begin(), v.end(), kout);
        return cudaSuccess;
    }
};

struct DeviceReduce {
    template <class K, class V, class R, class N>
    static cudaError_t ReduceByKey(void* tmp, size_t& tb, const K* kin, K* uniq, const V* vin, V* agg, R* nruns, Sum, N n,
                                   cudaStream_t = nullptr) {
        if (!tmp) {
            tb = 16;
            return cudaSuccess;
        }
        size_t r = 0;
        for (size_t i = 0; i < (size_t)n;) {
            size_t j = i;
            V s = 0;
            while (j < (size_t)n && kin[j] == kin[i]) s += vin[j++];
            uniq[r] = kin[i];
            agg[r] = s;
            r++;
            i = j;
        }
        *nruns = (R)r;
        return cudaSuccess;
    }
};

struct DeviceRunLengthEncode {
    template <class K, class C, class R, class N>
    static cudaError_t Encode(void* tmp, size_t& tb, const K* in, K* uniq, C* counts, R* nruns, N n, cudaStream_t = nullptr) {
        if (!tmp) {
            tb = 16;
            return cudaSuccess;
        }
        size_t r = 0;
        for (size_t i = 0; i < (size_t)n;) {
            size_t j = i;
            while (j < (size_t)n && in[j] == in[i]) j++;
            uniq[r] = in[i];
            counts[r] = (C)(j - i);
            r++;
            i = j;
        }
        *nruns = (R)r;
        return cudaSuccess;
    }
};
}  // namespace cub

// Shim for <Random123/array.h>: fixed-size counter/key arrays with the
// little-endian multi-word increment Random123 documents (r123arrayNxW::incr).
// Pinned by the carry tests the reference keeps at
// RandBLAS/test/basic_rng/test_r123.cc:735-796 (re-expressed in tests/test_oracle_rng.py).
#pragma once
#include "features/compilerfeatures.h"
#include <cstring>

template <typename T, int N>
struct r123array {
    using value_type = T;
    static constexpr int static_size = N;
    T v[N];

    T&       operator[](int i)       { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T*       begin()       { return v; }
    const T* begin() const { return v; }
    T*       end()         { return v + N; }
    const T* end()   const { return v + N; }
    T*       data()        { return v; }
    const T* data()  const { return v; }
    static constexpr size_t size() { return N; }

    // add an unsigned 64-bit step to the N*W-bit little-endian integer (wraps mod 2^(N*W))
    r123array& incr(R123_ULONG_LONG n = 1) {
        constexpr int W = 8 * sizeof(T);
        if constexpr (W >= 64) {
            T old = v[0];
            v[0] = (T)(v[0] + n);
            bool carry = v[0] < old;
            for (int i = 1; i < N && carry; ++i) { v[i] = (T)(v[i] + 1); carry = (v[i] == 0); }
        } else {
            unsigned long long carry = n;
            for (int i = 0; i < N && carry; ++i) {
                unsigned long long lo = carry & ((1ull << W) - 1);
                unsigned long long s  = (unsigned long long)v[i] + lo;
                v[i]  = (T)s;
                carry = (carry >> W) + (s >> W);
            }
        }
        return *this;
    }
    bool operator==(const r123array& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const r123array& o) const { return !(*this == o); }
};

using r123array2x32 = r123array<uint32_t, 2>;
using r123array4x32 = r123array<uint32_t, 4>;
using r123array2x64 = r123array<uint64_t, 2>;
using r123array4x64 = r123array<uint64_t, 4>;

// Shim for <Random123/philox.h>: Philox4x32-10 and Philox2x32-10 as published in
// Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11).
// Pinned by the known-answer vectors the reference ships at
// RandBLAS/test/basic_rng/r123_kat_vectors.txt:12-14,19-21.
#pragma once
#include "array.h"

namespace r123 {

template <int ROUNDS>
struct Philox4x32_R {
    using ctr_type  = r123array4x32;
    using key_type  = r123array2x32;
    using ukey_type = r123array2x32;
    static constexpr unsigned rounds = ROUNDS;
    ctr_type operator()(ctr_type c, key_type k) const {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        for (int r = 0; r < ROUNDS; ++r) {
            uint64_t p0 = (uint64_t)M0 * c.v[0];
            uint64_t p1 = (uint64_t)M1 * c.v[2];
            ctr_type o;
            o.v[0] = (uint32_t)(p1 >> 32) ^ c.v[1] ^ k.v[0];
            o.v[1] = (uint32_t)p1;
            o.v[2] = (uint32_t)(p0 >> 32) ^ c.v[3] ^ k.v[1];
            o.v[3] = (uint32_t)p0;
            c = o;
            k.v[0] += W0; k.v[1] += W1;
        }
        return c;
    }
};

template <int ROUNDS>
struct Philox2x32_R {
    using ctr_type  = r123array2x32;
    using key_type  = r123array<uint32_t, 1>;
    using ukey_type = key_type;
    static constexpr unsigned rounds = ROUNDS;
    ctr_type operator()(ctr_type c, key_type k) const {
        const uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
        for (int r = 0; r < ROUNDS; ++r) {
            uint64_t p = (uint64_t)M * c.v[0];
            ctr_type o;
            o.v[0] = (uint32_t)(p >> 32) ^ k.v[0] ^ c.v[1];
            o.v[1] = (uint32_t)p;
            c = o;
            k.v[0] += W;
        }
        return c;
    }
};

using Philox4x32 = Philox4x32_R<10>;
using Philox2x32 = Philox2x32_R<10>;

} // namespace r123

// On-chip Philox4x32-10 + Random123-convention Box-Muller (device).
//
// Follows the published Philox algorithm (Salmon et al., SC'11) that RandBLAS uses through
// r123::Philox4x32 (RandBLAS/RandBLAS/base.hh:53) and the Random123 u01 / uneg11 / boxmuller
// conventions RandBLAS applies in r123ext::boxmul / uneg11 (RandBLAS/RandBLAS/random_gen.hh:87-165).
// The integer stream is bit-exact with the reference.  The float Box-Muller values are evaluated in
// fp64 and rounded once to fp32, which reproduces a correctly-rounded host sinf/cosf/logf; the host
// libm is within 1 float ulp of that (tests/test_gpu_fill.py states the tolerance).
#pragma once
#include <cstdint>

namespace rlb {

struct Ctr128 {
    uint32_t v[4];
};

// 128-bit little-endian counter + 64-bit step (wraps), = r123array4x32::incr(n)
__host__ __device__ __forceinline__ Ctr128 ctr_add(Ctr128 c, uint64_t step) {
    uint64_t s = (uint64_t)c.v[0] + (step & 0xFFFFFFFFull);
    c.v[0] = (uint32_t)s;
    s = (uint64_t)c.v[1] + (step >> 32) + (s >> 32);
    c.v[1] = (uint32_t)s;
    s = (uint64_t)c.v[2] + (s >> 32);
    c.v[2] = (uint32_t)s;
    c.v[3] += (uint32_t)(s >> 32);
    return c;
}

__host__ __device__ __forceinline__ void philox4x32_10(const Ctr128& ctr, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c0 = ctr.v[0], c1 = ctr.v[1], c2 = ctr.v[2], c3 = ctr.v[3];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifdef __CUDACC__
__device__ __forceinline__ float u01f(uint32_t u) { return __fmaf_rn(__uint2float_rn(u), 0x1p-32f, 0x1p-33f); }
__device__ __forceinline__ float uneg11f(uint32_t u) { return __fmaf_rn(__int2float_rn((int32_t)u), 0x1p-31f, 0x1p-32f); }

// (r sin(pi x), r cos(pi x)), x = uneg11(u0), r = sqrt(-2 ln u01(u1)); the reference's host shim
// evaluates sin/cos at the float product PIf*x (random_gen.hh:56-60), reproduced here.
__device__ __forceinline__ void boxmuller_f(uint32_t u0, uint32_t u1, float& o0, float& o1) {
    const float PIf = 3.1415926535897932f;
    float t = __fmul_rn(PIf, uneg11f(u0));
    double sd, cd;
    sincos((double)t, &sd, &cd);
    float s = __double2float_rn(sd), c = __double2float_rn(cd);
    float l = __double2float_rn(log((double)u01f(u1)));
    float r = __fsqrt_rn(__fmul_rn(-2.f, l));
    o0 = __fmul_rn(s, r);
    o1 = __fmul_rn(c, r);
}

// 4 matrix entries for one counter: r123ext::boxmul::generate / uneg11::generate
template <int FAMILY>
__device__ __forceinline__ void generate4(const Ctr128& ctr, uint32_t k0, uint32_t k1, float rv[4]) {
    uint32_t r[4];
    philox4x32_10(ctr, k0, k1, r);
    if (FAMILY == RLB200_FAMILY_GAUSSIAN) {
        boxmuller_f(r[0], r[1], rv[0], rv[1]);
        boxmuller_f(r[2], r[3], rv[2], rv[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) rv[i] = uneg11f(r[i]);
    }
}
#endif

}  // namespace rlb

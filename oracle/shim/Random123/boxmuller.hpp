// Shim for <Random123/boxmuller.hpp>: Box-Muller on a pair of random integers,
// (r sin(pi*uneg11(u0)), r cos(pi*uneg11(u0))), r = sqrt(-2 ln u01(u1)).
// sincospif/sincospi are the unqualified host functions RandBLAS itself defines
// just before including this header (random_gen.hh:54-66).
#pragma once
#include "uniform.hpp"
#include <cmath>

namespace r123 {

struct float2  { float  x, y; };
struct double2 { double x, y; };

R123_CUDA_DEVICE R123_STATIC_INLINE float2 boxmuller(uint32_t u0, uint32_t u1) {
    float r;
    float2 f;
    sincospif(uneg11<float>(u0), &f.x, &f.y);
    r = sqrtf(-2.f * logf(u01<float>(u1)));
    f.x *= r;
    f.y *= r;
    return f;
}

R123_CUDA_DEVICE R123_STATIC_INLINE double2 boxmuller(uint64_t u0, uint64_t u1) {
    double r;
    double2 f;
    sincospi(uneg11<double>(u0), &f.x, &f.y);
    r = sqrt(-2. * log(u01<double>(u1)));
    f.x *= r;
    f.y *= r;
    return f;
}

} // namespace r123

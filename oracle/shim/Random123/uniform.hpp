// Shim for <Random123/uniform.hpp>: integer -> floating conversions with
// Random123's documented conventions (u01: (0,1], uneg11: [-1,1], both centred
// on the half-open bins so 0 is never produced).
#pragma once
#include "features/compilerfeatures.h"
#include <array>
#include <limits>
#include <type_traits>

namespace r123 {

template <typename Ftype, typename Itype>
R123_CUDA_DEVICE R123_STATIC_INLINE Ftype u01(Itype in) {
    using Utype = typename std::make_unsigned<Itype>::type;
    constexpr Ftype factor = Ftype(1.) / (Ftype(std::numeric_limits<Utype>::max()) + Ftype(1.));
    constexpr Ftype halffactor = Ftype(0.5) * factor;
    return Utype(in) * factor + halffactor;
}

template <typename Ftype, typename Itype>
R123_CUDA_DEVICE R123_STATIC_INLINE Ftype uneg11(Itype in) {
    using Stype = typename std::make_signed<Itype>::type;
    constexpr Ftype factor = Ftype(1.) / (Ftype(std::numeric_limits<Stype>::max()) + Ftype(1.));
    constexpr Ftype halffactor = Ftype(0.5) * factor;
    return Stype(in) * factor + halffactor;
}

template <typename Ftype, typename CollType>
static inline std::array<Ftype, CollType::static_size> u01all(CollType in) {
    std::array<Ftype, CollType::static_size> ret;
    for (int i = 0; i < CollType::static_size; ++i) ret[i] = u01<Ftype>(in[i]);
    return ret;
}

template <typename Ftype, typename CollType>
static inline std::array<Ftype, CollType::static_size> uneg11all(CollType in) {
    std::array<Ftype, CollType::static_size> ret;
    for (int i = 0; i < CollType::static_size; ++i) ret[i] = uneg11<Ftype>(in[i]);
    return ret;
}

} // namespace r123

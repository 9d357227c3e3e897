// Shim for <Random123/features/compilerfeatures.h>.
// TEST INFRASTRUCTURE ONLY (oracle build). Random123 (DEShawResearch) is an
// un-vendored, un-pinned dependency of the reference (install/install.sh:342);
// this directory restates the small part of its published API that RandBLAS
// touches (RandBLAS/RandBLAS/random_gen.hh:35-38,68-69).
#pragma once
#include <cstdint>
#include <cstddef>
#define R123_CUDA_DEVICE
#define R123_STATIC_INLINE static inline
#define R123_FORCE_INLINE(decl) decl
#define R123_CONSTEXPR constexpr
#define R123_ULONG_LONG unsigned long long

// Shim for LAPACK++'s <lapack/config.h> — TEST INFRASTRUCTURE ONLY (oracle build).
#pragma once
#include <cstdint>
#include <cstddef>
typedef int lapack_int;      // LP64 OpenBLAS bundled with scipy
typedef int lapack_logical;

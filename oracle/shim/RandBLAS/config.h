// Hand-written per RandBLAS/RandBLAS/config.h.in:24-56 ("if you want to use RandBLAS without CMake").
#pragma once
#define RandBLAS_FULL_VERSION "oracle-shim"
#define RandBLAS_VERSION_MAJOR 1
#define RandBLAS_VERSION_MINOR 0
#define RandBLAS_VERSION_PATCH 0
#define RandBLAS_COMMITS_SINCE_RELEASE 0
#define RandBLAS_COMMIT_HASH "04f2018a"
#define RandBLAS_HAS_OpenMP

// array_common.cuh -- shared pieces of the ArrayKernel kernels: opcodes, launch geometry, counter-based RNG.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "sstc_internal.h"

namespace sstc {

// Opcode values of org.shared.array.kernel.ArrayKernel (src/org/shared/array/kernel/ArrayKernel.java:45-278).
enum : int { RR_SUM = 0, RR_PROD = 1, RR_MAX = 2, RR_MIN = 3, RR_VAR = 4 };
enum : int { RI_MAX = 0, RI_MIN = 1, RI_ZERO = 2, RI_GZERO = 3, RI_LZERO = 4, RI_SORT = 5 };
enum : int { RD_SUM = 0, RD_PROD = 1 };
enum : int { RA_SUM = 0, RA_PROD = 1, RA_VAR = 2, RA_MAX = 3, RA_MIN = 4, RA_ENT = 5 };
enum : int { CA_SUM = 0, CA_PROD = 1 };
enum : int {
    RU_ADD = 0, RU_MUL = 1, RU_ABS = 2, RU_POW = 3, RU_EXP = 4, RU_RND = 5, RU_LOG = 6, RU_SQRT = 7, RU_SQR = 8,
    RU_INV = 9, RU_COS = 10, RU_SIN = 11, RU_ATAN = 12, RU_FILL = 13, RU_SHUFFLE = 14
};
enum : int {
    CU_ADD = 0, CU_MUL = 1, CU_EXP = 2, CU_RND = 3, CU_CONJ = 4, CU_COS = 5, CU_SIN = 6, CU_FILL = 7, CU_SHUFFLE = 8
};
enum : int { IU_ADD = 0, IU_MUL = 1, IU_FILL = 2, IU_SHUFFLE = 3 };
enum : int { RE_ADD = 0, RE_SUB = 1, RE_MUL = 2, RE_DIV = 3, RE_MAX = 4, RE_MIN = 5 };
enum : int { IE_ADD = 0, IE_SUB = 1, IE_MUL = 2, IE_MAX = 3, IE_MIN = 4 };
enum : int { CE_ADD = 0, CE_SUB = 1, CE_MUL = 2, CE_DIV = 3 };
enum : int { C_TO_R_ABS = 0, C_TO_R_RE = 1, C_TO_R_IM = 2 };
enum : int { R_TO_C_RE = 0, R_TO_C_IM = 1 };
enum : int { I_TO_R = 0 };

constexpr int kMaxDims = 16;  // ranks beyond this fail with a message (the reference's tests use <= 5)

// Streaming kernels: a grid of (SM count x kCtasPerSm) CTAs of 256 threads, grid-stride loops.
constexpr int kThreads = 256;
constexpr int kCtasPerSm = 8;

int sm_count();

inline int stream_grid(int64_t work_items, int items_per_thread = 1) {
    int64_t want = (work_items + (int64_t) kThreads * items_per_thread - 1) / ((int64_t) kThreads * items_per_thread);
    int64_t cap = (int64_t) sm_count() * kCtasPerSm;
    int64_t g = want < cap ? want : cap;
    return (int) (g < 1 ? 1 : g);
}

// Counter-based generator for RND / SHUFFLE: splitmix64 of (seed, call, element). The reference's providers
// draw from java.util.Random (Arithmetic.java:602-611) or libc rand() (ElementOps.cpp:695-700) and are only
// specified statistically (ArrayKernelTest.java:156-172); this one is reproducible across launches.
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

__device__ inline double uniform01(uint64_t key, uint64_t index) {
    return (double) (splitmix64(key ^ splitmix64(index)) >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
}

// Stable ascending sort of (key, payload) pairs by (key, payload) inside every aligned segment of `seg` pairs
// (a power of two; `total` a multiple of it): the bitonic network of array_dimension.cu.
void sort_pairs_segments(uint64_t *keys, int32_t *pay, int64_t total, int64_t seg, cudaStream_t s);

uint64_t next_rng_key();  // advances the library-global call counter (array_elementwise.cu)

// Java Math.max / Math.min semantics (NaN wins, -0.0 < +0.0): ElementOps.java RE_MAX / RE_MIN go through
// Math.max/min; the reference's C++ uses std::max/min, which differ only on NaN and signed zeros.
// fmax/fmin (PTX max.f64 / min.f64) already order -0.0 below +0.0; they return the non-NaN operand, so the NaN
// case is patched with selects (no branches: these sit in the inner loops of the MAX / MIN folds).
__device__ __forceinline__ double java_max(double a, double b) {
    const double r = fmax(a, b);
    return (a != a) ? a : ((b != b) ? b : r);
}

__device__ __forceinline__ double java_min(double a, double b) {
    const double r = fmin(a, b);
    return (a != a) ? a : ((b != b) ? b : r);
}

// Single-pass 2-D inclusive sum scan of a dense row-major (rows x cols) array (array_scan2d.cu). `border`: the input is the
// destination with its interior replaced by the source shifted by (1, 1), which is ImageKernel.createIntegralImage.
bool scan2d_applicable(int64_t rows, int64_t cols);
void scan2d_sum(const double *src, int64_t spitch, double *dst, int64_t dpitch, int64_t rows, int64_t cols, bool border,
        cudaStream_t stream);

}  // namespace sstc

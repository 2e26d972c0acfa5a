// array_matrix.cu -- ArrayKernel.mul: row-major FP64 matrix product, real or interleaved complex.
//
// Replaces src/org/shared/array/kernel/MatrixOps.java:44-98 and native/src/shared/MatrixOps.cpp:26-98 (a naive
// i-j-k loop whose right operand is walked by column). This is the one real contraction on the path and the
// only FP64-compute-bound kernel: a shared-memory tiled, register-blocked kernel on the FP64 FMA pipe.
// (B200's tensor cores add nothing here: its FP64 tensor rate equals its FP64 FMA rate, and tcgen05 has no
// FP64 kind.) Every output element accumulates its k terms in ascending order in ONE accumulator, exactly
// like the reference's `sum = sum + l*r`, except that each step is a fused multiply-add.
//
// Real:    CTA tile 128 x 128, k-tile 8, 256 threads, 8 x 8 outputs per thread (rows ty + 16 i,
//          columns 2 tx + 32 g + e), operands staged through shared memory (A transposed to k-major).
// Complex: CTA tile 64 x 64, k-tile 8, 256 threads, 4 x 4 complex outputs per thread.
#include "array_common.cuh"

namespace sstc {

constexpr int BM = 128, BN = 128, BK = 8;  // 2 x (8 x 132 + 8 x 128) doubles = 33 KB of static shared memory

__global__ void __launch_bounds__(256) dgemm_kernel(const double *__restrict__ A, const double *__restrict__ B,
        double *__restrict__ C, int M, int N, int K) {
    __shared__ double As[2][BK][BM + 4];  // k-major so that a thread's 8 rows are a broadcast-friendly column
    __shared__ __align__(16) double Bs[2][BK][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            acc[i][j] = 0.0;
        }
    }

    // global -> register staging: A tile 128 x 8 and B tile 8 x 128, 4 values per thread each
    const int a_row = tid >> 1, a_k = (tid & 1) * 4;  // 128 rows x 2 halves of 4 consecutive k
    const int b_k = tid >> 5, b_col = (tid & 31) * 4;  // 8 k x 32 groups of 4 consecutive columns
    double ra[4], rb[4];

    auto load_tiles = [&](int k0) {
        const int gr = row0 + a_row;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int gk = k0 + a_k + e;
            ra[e] = (gr < M && gk < K) ? A[(int64_t) gr * K + gk] : 0.0;
        }
        const int gk = k0 + b_k;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int gc = col0 + b_col + e;
            rb[e] = (gk < K && gc < N) ? B[(int64_t) gk * N + gc] : 0.0;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            As[buf][a_k + e][a_row] = ra[e];
            Bs[buf][b_k][b_col + e] = rb[e];
        }
    };

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int ktiles = (K + BK - 1) / BK;
    for (int kt = 0; kt < ktiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) {
            load_tiles((kt + 1) * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; k++) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = As[buf][k][ty + 16 * i];
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const double2 v = *reinterpret_cast<const double2 *>(&Bs[buf][k][2 * tx + 32 * g]);
                b[2 * g] = v.x;
                b[2 * g + 1] = v.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc[i][j] = fma(a[i], b[j], acc[i][j]);
                }
            }
        }
        if (kt + 1 < ktiles) {
            store_tiles(buf ^ 1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int gr = row0 + ty + 16 * i;
        if (gr >= M) {
            continue;
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int gc = col0 + 2 * tx + 32 * g + e;
                if (gc < N) {
                    C[(int64_t) gr * N + gc] = acc[i][2 * g + e];
                }
            }
        }
    }
}

constexpr int ZM = 64, ZN = 64, ZK = 8;

__global__ void __launch_bounds__(256) zgemm_kernel(const double2 *__restrict__ A, const double2 *__restrict__ B,
        double2 *__restrict__ C, int M, int N, int K) {
    __shared__ double2 As[2][ZK][ZM + 2];
    __shared__ double2 Bs[2][ZK][ZN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.y * ZM, col0 = blockIdx.x * ZN;
    double2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc[i][j] = make_double2(0.0, 0.0);
        }
    }
    const int a_row = tid >> 2, a_k = (tid & 3) * 2;  // 64 rows x 4 groups of 2 k
    const int b_k = tid >> 5, b_col = (tid & 31) * 2;  // 8 k x 32 groups of 2 columns
    double2 ra[2], rb[2];
    auto load_tiles = [&](int k0) {
        const int gr = row0 + a_row;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int gk = k0 + a_k + e;
            ra[e] = (gr < M && gk < K) ? A[(int64_t) gr * K + gk] : make_double2(0.0, 0.0);
        }
        const int gk = k0 + b_k;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int gc = col0 + b_col + e;
            rb[e] = (gk < K && gc < N) ? B[(int64_t) gk * N + gc] : make_double2(0.0, 0.0);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            As[buf][a_k + e][a_row] = ra[e];
            Bs[buf][b_k][b_col + e] = rb[e];
        }
    };
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int ktiles = (K + ZK - 1) / ZK;
    for (int kt = 0; kt < ktiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) {
            load_tiles((kt + 1) * ZK);
        }
#pragma unroll
        for (int k = 0; k < ZK; k++) {
            double2 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a[i] = As[buf][k][ty + 16 * i];
                b[i] = Bs[buf][k][tx + 16 * i];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // sum + a*b with a*b = (ar br - ai bi, ar bi + ai br)  (Common.hpp:339-341)
                    acc[i][j].x = fma(a[i].x, b[j].x, fma(-a[i].y, b[j].y, acc[i][j].x));
                    acc[i][j].y = fma(a[i].x, b[j].y, fma(a[i].y, b[j].x, acc[i][j].y));
                }
            }
        }
        if (kt + 1 < ktiles) {
            store_tiles(buf ^ 1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int gr = row0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gc = col0 + tx + 16 * j;
            if (gr < M && gc < N) {
                C[(int64_t) gr * N + gc] = acc[i][j];
            }
        }
    }
}

__global__ void zero_kernel(double *dst, int64_t n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        dst[i] = 0.0;
    }
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_mul_dev(const double *lhs, int64_t nl, const double *rhs, int64_t nr, int lr, int rc, double *dst, int64_t ndst,
        int complex, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        // MatrixOps.cpp:60-74: the inner dimension is inferred from the array lengths
        const int64_t f = complex ? 2 : 1;
        if (nl < 0 || nr < 0 || ndst < 0) {
            throw std::runtime_error("Invalid arguments");
        }
        const int64_t lc = lr ? nl / (f * lr) : 0;
        const int64_t rr = rc ? nr / (f * rc) : 0;
        const int64_t inner = lc;
        if (lr < 0 || rc < 0 || nl != f * lr * lc || nr != f * rr * rc || ndst != f * (int64_t) lr * rc || inner != rr) {
            throw std::runtime_error("Invalid array lengths");
        }
        if ((nl > 0 && !lhs) || (nr > 0 && !rhs) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        if (lr == 0 || rc == 0) {
            return;
        }
        if (inner == 0) {
            zero_kernel<<<stream_grid(ndst), kThreads, 0, s>>>(dst, ndst);
            check_launch("zero_kernel");
            return;
        }
        if (inner > 2147483647LL) {
            throw std::runtime_error("Invalid array lengths");
        }
        if (complex) {
            if ((reinterpret_cast<uintptr_t>(lhs) | reinterpret_cast<uintptr_t>(rhs) | reinterpret_cast<uintptr_t>(dst)) & 15) {
                throw std::runtime_error("complex mul needs 16-byte aligned device arrays");
            }
            dim3 grid((unsigned) ((rc + ZN - 1) / ZN), (unsigned) ((lr + ZM - 1) / ZM));
            zgemm_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const double2 *>(lhs), reinterpret_cast<const double2 *>(rhs),
                    reinterpret_cast<double2 *>(dst), lr, rc, (int) inner);
        } else {
            dim3 grid((unsigned) ((rc + BN - 1) / BN), (unsigned) ((lr + BM - 1) / BM));
            dgemm_kernel<<<grid, 256, 0, s>>>(lhs, rhs, dst, lr, rc, (int) inner);
        }
        check_launch("gemm_kernel");
    });
}

}  // extern "C"

// array_matrix.cu -- ArrayKernel.mul: row-major FP64 matrix product, real or interleaved complex.
//
// Replaces src/org/shared/array/kernel/MatrixOps.java:44-98 and native/src/shared/MatrixOps.cpp:26-98 (a naive
// i-j-k loop whose right operand is walked by column). This is the one real contraction on the path and the
// only FP64-compute-bound kernel. tcgen05 has no FP64 kind, so it runs on the FP64 tensor pipe through
// mma.sync m8n8k4 (SASS DMMA.8x8x4, the only FP64 MMA shape sm_100a executes natively; the larger PTX shapes are
// split into it). Measured at 4096^2 on B200: 32.2 TFLOP/s real, 26.8 complex (cuBLAS DGEMM on the same box:
// 35.4); the first version of this file, a 128 x 128 x 8 register-blocked kernel on the FP64 FMA pipe, reached
// 17.1. Sums are accumulated in k-chunks of 4 in ascending order; the reference adds one term at a time, so
// results agree to rounding (relative L2 <= 1e-12 log2 K in the tests), not bit for bit.
#include <stdlib.h>

#include <mutex>

#include "array_common.cuh"

namespace sstc {

// ---- FP64 tensor-core path (DMMA.8x8x4, the native FP64 MMA shape of sm_100a) -------------------------------
// CTA tile 128 x 128, k-tile 16, 8 warps as 2 (rows) x 4 (columns), warp tile 64 x 32 = 8 x 4 MMA tiles,
// 64 accumulator doubles per thread. Operands go global -> shared with cp.async (16-byte chunks when the rows are
// 16-byte aligned, else 8-byte), three stages deep, one CTA barrier per k-tile. Shared-memory pitches (20 and 132
// doubles) are = 4 mod 16, which makes every fragment load (8 rows x 4 k, 64-bit) conflict-free.
//
// Complex (interleaved) products run on the same kernel: C(M x 2N real) = A(M x 2K real) * B', where row 2k of B'
// is row k of B as stored (br, bi) and row 2k+1 is (-bi, br). B' is never materialised: the B fragment for an
// odd real k reads the raw row at column n ^ 1 and negates it on even n.
constexpr int TM = 128, TN = 128, TK = 16, kStages = 3;
constexpr int A_LD = TK + 4, B_LD = TN + 4;
constexpr int kStageDoubles = TM * A_LD + TK * B_LD;

__device__ __forceinline__ void cp_async(double *smem_dst, const double *gsrc, bool valid, int vec) {
    const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
    const int src = valid ? vec * 8 : 0;  // src-size 0: the destination is zero-filled, the source is not read
    if (vec == 2) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src) : "memory");
    }
}

// A: M x K row-major (complex: K counts real columns, i.e. 2 x the complex inner dimension).
// B: real: K x N row-major; complex: K/2 raw rows x N (N counts real columns).
template <int VEC, bool COMPLEX>
__global__ void __launch_bounds__(256, 1) dmma_gemm_kernel(const double *__restrict__ A, const double *__restrict__ B,
        double *__restrict__ C, int M, int N, int K) {
    extern __shared__ __align__(16) double gsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    constexpr int BROWS = COMPLEX ? TK / 2 : TK;  // raw rows of B per k-tile
    const int Kb = COMPLEX ? K / 2 : K;           // raw rows of B

    auto load_tile = [&](int stage, int k0) {
        double *As = gsm + (size_t) stage * kStageDoubles, *Bs = As + TM * A_LD;
        constexpr int ACH = TK / VEC;  // chunks per row of the A tile
#pragma unroll
        for (int i = 0; i < TM * ACH / 256; i++) {
            const int id = tid + 256 * i, r = id / ACH, c = (id % ACH) * VEC;
            const bool ok = row0 + r < M && k0 + c < K;
            cp_async(As + r * A_LD + c, ok ? A + (int64_t) (row0 + r) * K + k0 + c : A, ok, VEC);
        }
        constexpr int BCH = TN / VEC;
        const int kb0 = COMPLEX ? k0 / 2 : k0;
#pragma unroll
        for (int i = 0; i < BROWS * BCH / 256; i++) {
            const int id = tid + 256 * i, r = id / BCH, c = (id % BCH) * VEC;
            const bool ok = kb0 + r < Kb && col0 + c < N;
            cp_async(Bs + r * B_LD + c, ok ? B + (int64_t) (kb0 + r) * N + col0 + c : B, ok, VEC);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc[i][j][0] = 0.0;
            acc[i][j][1] = 0.0;
        }
    }
    const int ktiles = (K + TK - 1) / TK;
#pragma unroll
    for (int st = 0; st < kStages - 1; st++) {
        if (st < ktiles) {
            load_tile(st, st * TK);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kt = 0; kt < ktiles; kt++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
        __syncthreads();  // tile kt has landed for every thread; the stage refilled below was drained in iteration kt-1
        if (kt + kStages - 1 < ktiles) {
            load_tile((kt + kStages - 1) % kStages, (kt + kStages - 1) * TK);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double *As = gsm + (size_t) (kt % kStages) * kStageDoubles, *Bs = As + TM * A_LD;
#pragma unroll
        for (int s = 0; s < TK / 4; s++) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a[i] = As[(wm + i * 8 + g) * A_LD + s * 4 + tig];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = wn + j * 8 + g;
                if (COMPLEX) {
                    const int kk = s * 4 + tig, p = kk & 1;
                    const double v = Bs[(kk >> 1) * B_LD + (n ^ p)];
                    b[j] = (p && !(g & 1)) ? -v : v;
                } else {
                    b[j] = Bs[(s * 4 + tig) * B_LD + n];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                 : "d"(a[i]), "d"(b[j]));
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int gr = row0 + wm + i * 8 + g;
        if (gr >= M) {
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gc = col0 + wn + j * 8 + 2 * tig;
            double *dst = C + (int64_t) gr * N + gc;
            if (VEC == 2) {  // N even, C 16-byte aligned: the pair is inside or outside together
                if (gc < N) {
                    *reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            } else {
                if (gc < N) {
                    dst[0] = acc[i][j][0];
                }
                if (gc + 1 < N) {
                    dst[1] = acc[i][j][1];
                }
            }
        }
    }
}

template <int VEC, bool COMPLEX>
static void launch_dmma(const double *A, const double *B, double *C, int M, int N, int K, cudaStream_t s) {
    constexpr size_t smem = (size_t) kStages * kStageDoubles * sizeof(double);
    static std::once_flag once[16];
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::call_once(once[dev & 15], [] {
        SSTC_CUDA(cudaFuncSetAttribute(dmma_gemm_kernel<VEC, COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                (int) smem));
    });
    dim3 grid((unsigned) ((N + TN - 1) / TN), (unsigned) ((M + TM - 1) / TM));
    dmma_gemm_kernel<VEC, COMPLEX><<<grid, 256, smem, s>>>(A, B, C, M, N, K);
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
        const bool aligned = !((reinterpret_cast<uintptr_t>(lhs) | reinterpret_cast<uintptr_t>(rhs) | reinterpret_cast<uintptr_t>(dst)) & 15);
        if (complex) {
            if (!aligned) {
                throw std::runtime_error("complex mul needs 16-byte aligned device arrays");
            }
            if ((int64_t) 2 * inner > 2147483647LL || (int64_t) 2 * rc > 2147483647LL) {
                throw std::runtime_error("Invalid array lengths");
            }
            launch_dmma<2, true>(lhs, rhs, dst, lr, 2 * rc, 2 * (int) inner, s);
        } else if (aligned && inner % 2 == 0 && rc % 2 == 0) {
            launch_dmma<2, false>(lhs, rhs, dst, lr, rc, (int) inner, s);
        } else {
            launch_dmma<1, false>(lhs, rhs, dst, lr, rc, (int) inner, s);
        }
        check_launch("gemm_kernel");
    });
}

}  // extern "C"

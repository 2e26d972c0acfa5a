// fft_aux_kernels.cuh -- helper kernels that extend the axis pass to every length:
//   * four-step decomposition for powers of two above 4096 (twiddle + transpose between two sub-passes),
//   * Bluestein's chirp-z for lengths the shared-memory kernels cannot hold (any length, via power-of-two
//     transforms), and
//   * real <-> complex staging so that r2c / c2r can use either of them on the last axis.
// The reference handles any length with one recursive routine (FftOps.java:173-248, O(n^2) for primes); these
// keep the provider a drop-in for any length while the arithmetic stays on the GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sstc {

// exp(-+ 2 pi i num / den) with an exact integer phase (sincospi reduces the argument exactly)
__device__ __forceinline__ double2 unit_root_dev(int64_t num, int64_t den, int inverse) {
    double s, c;
    sincospi(2.0 * (double) num / (double) den, &s, &c);
    return make_double2(c, inverse ? s : -s);
}

__device__ __forceinline__ double2 cmul_dev(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Four-step middle step: A is (outer, L1, L2, inner) after the length-L1 sub-pass; T is (outer, L2, L1, inner):
// T[o][n2][k1][i] = A[o][k1][n2][i] * w_L^{n2*k1}.  inner == 1 goes through 32x32 shared-memory tiles.
__global__ void __launch_bounds__(256) fourstep_twiddle_transpose_tile_kernel(const double2 *__restrict__ A, double2 *T,
        int64_t outer, int L1, int L2, int inverse) {
    __shared__ double2 tile[32][33];
    const int64_t L = (int64_t) L1 * L2;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t1 = (L1 + 31) / 32, t2 = (L2 + 31) / 32;
    const int64_t ntiles = outer * t1 * t2;
    for (int64_t blk = blockIdx.x; blk < ntiles; blk += gridDim.x) {
        const int i2 = (int) (blk % t2);
        const int i1 = (int) ((blk / t2) % t1);
        const int64_t o = blk / ((int64_t) t1 * t2);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int k1 = i1 * 32 + ty + 8 * r, n2 = i2 * 32 + tx;
            if (k1 < L1 && n2 < L2) {
                const double2 w = unit_root_dev(((int64_t) n2 * k1) % L, L, inverse);
                tile[ty + 8 * r][tx] = cmul_dev(A[o * L + (int64_t) k1 * L2 + n2], w);
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int n2 = i2 * 32 + ty + 8 * r, k1 = i1 * 32 + tx;
            if (k1 < L1 && n2 < L2) {
                T[o * L + (int64_t) n2 * L1 + k1] = tile[tx][ty + 8 * r];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) fourstep_twiddle_transpose_kernel(const double2 *__restrict__ A, double2 *T,
        int64_t outer, int L1, int L2, int64_t inner, int inverse) {
    const int64_t L = (int64_t) L1 * L2;
    const int64_t total = outer * L * inner;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e % inner;
        int64_t rem = e / inner;
        const int k1 = (int) (rem % L1);  // destination order (o, n2, k1, i): coalesced stores
        rem /= L1;
        const int n2 = (int) (rem % L2);
        const int64_t o = rem / L2;
        const double2 w = unit_root_dev(((int64_t) n2 * k1) % L, L, inverse);
        T[e] = cmul_dev(A[((o * L1 + k1) * L2 + n2) * inner + i], w);
    }
}

// ---- Bluestein ---------------------------------------------------------------------------------------------
// chirp[n] = exp(-+ pi i n^2 / L); the phase n^2 mod 2L is exact in 64-bit integers.
__device__ __forceinline__ double2 chirp_dev(int64_t n, int64_t L, int inverse) {
    const int64_t ph = (n * n) % (2 * L);
    double s, c;
    sincospi((double) ph / (double) L, &s, &c);
    return make_double2(c, inverse ? s : -s);
}

// B[m], m in [0, M): the wrapped conjugate chirp, b[m] = conj(chirp[m]) for m < L, b[M - m] = b[m], else 0
__global__ void __launch_bounds__(256) bluestein_filter_kernel(double2 *B, int64_t L, int64_t M, int inverse) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
        double2 v = make_double2(0.0, 0.0);
        if (m < L) {
            const double2 c = chirp_dev(m, L, inverse);
            v = make_double2(c.x, -c.y);
        } else if (M - m < L) {
            const double2 c = chirp_dev(M - m, L, inverse);
            v = make_double2(c.x, -c.y);
        }
        B[m] = v;
    }
}

// Y[line][m] = x[line][m] * chirp[m] for m < L, 0 for L <= m < M.  Lines are gathered from the
// (outer, L, inner) view: line = o*inner + i.
__global__ void __launch_bounds__(256) bluestein_pre_kernel(const double2 *__restrict__ x, double2 *Y, int64_t outer,
        int64_t L, int64_t inner, int64_t M, int inverse) {
    const int64_t total = outer * inner * M;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t m = e % M, line = e / M;
        double2 v = make_double2(0.0, 0.0);
        if (m < L) {
            const int64_t o = line / inner, i = line - o * inner;
            v = cmul_dev(x[(o * L + m) * inner + i], chirp_dev(m, L, inverse));
        }
        Y[e] = v;
    }
}

__global__ void __launch_bounds__(256) bluestein_mul_kernel(double2 *Y, const double2 *__restrict__ Bf, int64_t lines,
        int64_t M) {
    const int64_t total = lines * M;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        Y[e] = cmul_dev(Y[e], Bf[e % M]);
    }
}

// out[(o*L + k)*inner + i] = Z[line][k] * chirp[k] * scale   (Z already carries the 1/M of the inverse pass)
__global__ void __launch_bounds__(256) bluestein_post_kernel(const double2 *__restrict__ Z, double2 *out, int64_t outer,
        int64_t L, int64_t inner, int64_t M, int inverse, double scale) {
    const int64_t total = outer * L * inner;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e % inner;
        const int64_t rem = e / inner;
        const int64_t k = rem % L, o = rem / L;
        const double2 v = cmul_dev(Z[(o * inner + i) * M + k], chirp_dev(k, L, inverse));
        out[e] = make_double2(v.x * scale, v.y * scale);
    }
}

// ---- real <-> complex staging for the last axis ---------------------------------------------------------------
__global__ void __launch_bounds__(256) real_to_complex_kernel(const double *__restrict__ in, double2 *out, int64_t n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        out[e] = make_double2(in[e], 0.0);
    }
}

// half[line][k] = full[line][k], k in [0, L/2]
__global__ void __launch_bounds__(256) truncate_half_kernel(const double2 *__restrict__ full, double2 *half, int64_t lines,
        int64_t L) {
    const int64_t h = L / 2 + 1;
    const int64_t total = lines * h;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t line = e / h, k = e - line * h;
        half[e] = full[line * L + k];
    }
}

// full[line][k] = half[line][k] for k <= L/2, conj(half[line][L-k]) otherwise
__global__ void __launch_bounds__(256) hermitian_expand_kernel(const double2 *__restrict__ half, double2 *full,
        int64_t lines, int64_t L) {
    const int64_t h = L / 2 + 1;
    const int64_t total = lines * L;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t line = e / L, k = e - line * L;
        double2 v;
        if (k < h) {
            v = half[line * h + k];
        } else {
            v = half[line * h + (L - k)];
            v.y = -v.y;
        }
        full[e] = v;
    }
}

__global__ void __launch_bounds__(256) complex_to_real_kernel(const double2 *__restrict__ in, double *out, int64_t n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        out[e] = in[e].x;
    }
}

}  // namespace sstc

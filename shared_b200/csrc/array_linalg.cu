// array_linalg.cu -- ArrayKernel.invert and ArrayKernel.find.
//
// invert replaces native/src/shared/LinearAlgebraOpsLu.cpp:26-176 (JAMA's LU with partial pivoting, then L / U
// solves against the permuted identity; Java twin: LinearAlgebraOps.java). Same factorisation, same pivot rule
// (largest magnitude in the column, the first one on ties), run right-looking on the GPU: per column one small
// kernel (pivot search, row swap, multipliers) and one grid-wide rank-1 update; the two triangular solves are a
// rank-1 update per column as well. Rounding differs from the reference's left-looking dot products in the last
// bits only. A zero on U's diagonal fails with "Matrix is singular" (LinearAlgebraOpsLu.cpp:62-64).
// This is a correctness-first kernel set (5n launches, every update streams the matrix): invert is not on the
// north-star hot path; it is here so that callers of Matrix.mInvert() keep working under the CUDA provider.
//
// find replaces native/src/shared/IndexOps.cpp:26-130 (IndexOps.java:44-91): along one line of an int array (as
// written by riOp: positions, then -1 padding) count the non-negative entries and return that many leading entries.
#include <math.h>

#include <string>
#include <vector>

#include "array_common.cuh"

namespace sstc {

// One CTA. Column j of the row-major n x n matrix: pivot = first row of largest magnitude at or below the diagonal,
// swapped into place (whole rows + permutation), then the multipliers below the diagonal.
__global__ void __launch_bounds__(256) lu_pivot_kernel(double *lu, int32_t *piv, int n, int j) {
    __shared__ double sval[256];
    __shared__ int sidx[256];
    const int tid = threadIdx.x;
    double best = -1.0;
    int bi = j;
    for (int i = j + tid; i < n; i += 256) {
        const double v = fabs(lu[(int64_t) i * n + j]);
        if (v > best) {
            best = v;
            bi = i;
        }
    }
    sval[tid] = best;
    sidx[tid] = bi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) {
            const double v = sval[tid + w];
            const int i = sidx[tid + w];
            if (v > sval[tid] || (v == sval[tid] && i < sidx[tid])) {
                sval[tid] = v;
                sidx[tid] = i;
            }
        }
        __syncthreads();
    }
    const int p = sidx[0];
    if (p != j) {
        for (int k = tid; k < n; k += 256) {
            const double t = lu[(int64_t) p * n + k];
            lu[(int64_t) p * n + k] = lu[(int64_t) j * n + k];
            lu[(int64_t) j * n + k] = t;
        }
        if (tid == 0) {
            const int32_t t = piv[p];
            piv[p] = piv[j];
            piv[j] = t;
        }
    }
    __syncthreads();
    const double d = lu[(int64_t) j * n + j];
    if (d != 0.0) {
        for (int i = j + 1 + tid; i < n; i += 256) {
            lu[(int64_t) i * n + j] /= d;
        }
    }
}

// a[i][k] -= a[i][j] * a[j][k] for i, k > j.
__global__ void __launch_bounds__(256) lu_update_kernel(double *lu, int n, int j) {
    const int k = j + 1 + blockIdx.x * 32 + (threadIdx.x & 31);
    const int i0 = j + 1 + blockIdx.y * 64 + (threadIdx.x >> 5);
    if (k >= n) {
        return;
    }
    const double u = lu[(int64_t) j * n + k];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int i = i0 + 8 * r;
        if (i < n) {
            lu[(int64_t) i * n + k] -= lu[(int64_t) i * n + j] * u;
        }
    }
}

__global__ void __launch_bounds__(256) permuted_identity_kernel(double *b, const int32_t *piv, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        b[(int64_t) i * n + piv[i]] = 1.0;  // LinearAlgebraOpsLu.cpp:66
    }
}

// b[i][c] -= b[k][c] * lu[i][k] for the rows i in (k, n) (LOWER) or [0, k) (!LOWER), every column c.
template <bool LOWER>
__global__ void __launch_bounds__(256) tri_update_kernel(const double *lu, double *b, int n, int k) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i0 = (LOWER ? k + 1 : 0) + blockIdx.y * 64 + (threadIdx.x >> 5);
    const int iend = LOWER ? n : k;
    if (c >= n) {
        return;
    }
    const double bk = b[(int64_t) k * n + c];
    if (bk == 0.0) {
        return;  // exact: subtracting 0 * l changes nothing (finite l); skips the identity's zero columns early on
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int i = i0 + 8 * r;
        if (i < iend) {
            b[(int64_t) i * n + c] -= bk * lu[(int64_t) i * n + k];
        }
    }
}

__global__ void __launch_bounds__(256) scale_row_kernel(const double *lu, double *b, int n, int k) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) {
        b[(int64_t) k * n + c] /= lu[(int64_t) k * n + k];
    }
}

__global__ void __launch_bounds__(256) diag_zero_kernel(const double *lu, int n, int *flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && lu[(int64_t) i * n + i] == 0.0) {
        *flag = 1;
    }
}

__global__ void __launch_bounds__(256) iota_kernel(int32_t *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        p[i] = i;
    }
}

// One CTA: out[i] = line[i * stride] for the whole line, *count = number of non-negative entries.
__global__ void __launch_bounds__(256) find_kernel(const int32_t *__restrict__ src, int64_t offset, int64_t stride, int size,
        int32_t *out, int32_t *count) {
    __shared__ int part[256];
    int c = 0;
    for (int i = threadIdx.x; i < size; i += 256) {
        const int32_t v = src[offset + (int64_t) i * stride];
        out[i] = v;
        c += v >= 0;
    }
    part[threadIdx.x] = c;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            part[threadIdx.x] += part[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *count = part[0];
    }
}

static thread_local Scratch t_lu, t_piv, t_small;

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_invert_dev(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        const int64_t n2 = (int64_t) size * size;
        if (size < 0 || nsrc != n2 || ndst != n2 || (n2 > 0 && (!src || !dst))) {
            throw std::runtime_error("Invalid arguments");  // LinearAlgebraOpsLu.cpp:31-40
        }
        if (size == 0) {
            return;
        }
        const int n = size;
        double *lu = static_cast<double *>(t_lu.get(n2 * 8));
        int32_t *piv = static_cast<int32_t *>(t_piv.get((int64_t) n * 4));
        int *flag = static_cast<int *>(t_small.get(sizeof(int)));
        SSTC_CUDA(cudaMemcpyAsync(lu, src, (size_t) n2 * 8, cudaMemcpyDeviceToDevice, s));
        SSTC_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
        const int g1 = (n + 255) / 256;
        iota_kernel<<<g1, 256, 0, s>>>(piv, n);
        for (int j = 0; j < n; j++) {
            lu_pivot_kernel<<<1, 256, 0, s>>>(lu, piv, n, j);
            const int rem = n - j - 1;
            if (rem > 0) {
                lu_update_kernel<<<dim3((unsigned) ((rem + 31) / 32), (unsigned) ((rem + 63) / 64)), 256, 0, s>>>(lu, n, j);
            }
        }
        check_launch("lu kernels");
        count_launch(2 * n - 1);
        diag_zero_kernel<<<g1, 256, 0, s>>>(lu, n, flag);
        int singular = 0;
        SSTC_CUDA(cudaMemcpyAsync(&singular, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
        if (singular) {
            throw std::runtime_error("Matrix is singular");
        }
        // dst = P I (the reference assumes a zero-filled dst, LinearAlgebraOpsLu.cpp:66; here it is cleared), then
        // L Y = P I and U X = Y in place
        SSTC_CUDA(cudaMemsetAsync(dst, 0, (size_t) n2 * 8, s));
        permuted_identity_kernel<<<g1, 256, 0, s>>>(dst, piv, n);
        const unsigned gc = (unsigned) ((n + 31) / 32);
        for (int k = 0; k < n - 1; k++) {
            const int rows = n - k - 1;
            tri_update_kernel<true><<<dim3(gc, (unsigned) ((rows + 63) / 64)), 256, 0, s>>>(lu, dst, n, k);
        }
        for (int k = n - 1; k >= 0; k--) {
            scale_row_kernel<<<g1, 256, 0, s>>>(lu, dst, n, k);
            if (k > 0) {
                tri_update_kernel<false><<<dim3(gc, (unsigned) ((k + 63) / 64)), 256, 0, s>>>(lu, dst, n, k);
            }
        }
        check_launch("lu solve kernels");
        count_launch(3 * n);
    });
}

int sstc_find_dev(const int32_t *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, const int32_t *logical,
        int32_t *out, int32_t *count, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        if (!sD || !sS || !logical || !count || nd < 0 || nsrc < 0 || (nsrc > 0 && !src)) {
            throw std::runtime_error("Invalid arguments");
        }
        int64_t acc = 0;
        for (int i = 0; i < nd; i++) {  // MappingOps::checkDimensions
            if (sD[i] < 0 || sS[i] < 0) {
                throw std::runtime_error("Invalid dimensions and/or strides");
            }
            acc += (int64_t) (sD[i] - 1) * sS[i];
        }
        if (acc != nsrc - 1) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
        int active = -1, markers = 0;
        for (int d = 0; d < nd; d++) {
            if (logical[d] == -1) {
                active = d;
                markers++;
            }
        }
        if (markers != 1) {
            throw std::runtime_error("Invalid arguments");  // IndexOps.cpp:98-100 (Java: "The dimension marker must be unique")
        }
        int64_t offset = 0;
        for (int d = 0; d < nd; d++) {
            if (d != active) {
                if (!(logical[d] >= 0 && logical[d] < sD[d])) {
                    throw std::runtime_error("Invalid index");
                }
                offset += (int64_t) logical[d] * sS[d];
            }
        }
        const int size = sD[active];
        *count = 0;
        if (size == 0) {
            return;
        }
        if (!out) {
            throw std::runtime_error("Invalid arguments");
        }
        int32_t *dcount = static_cast<int32_t *>(t_small.get(sizeof(int32_t)));
        find_kernel<<<1, 256, 0, s>>>(src, offset, sS[active], size, out, dcount);
        check_launch("find_kernel");
        SSTC_CUDA(cudaMemcpyAsync(count, dcount, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
    });
}

}  // extern "C"

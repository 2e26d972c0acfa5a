// array_linalg.cu -- ArrayKernel.invert, ArrayKernel.svd and ArrayKernel.find.
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

#include <algorithm>
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

// ---- svd: one-sided Jacobi (Hestenes) ---------------------------------------------------------------------------
// The reference runs JAMA's Golub-Kahan bidiagonalisation + implicit QR (LinearAlgebraOpsSvd.cpp:26-560), a serial
// algorithm. Here: rotate column pairs of a column-major working copy W until all columns are orthogonal, the same
// rotations accumulated into V; then sigma_j = |W_j|, U_j = W_j / sigma_j, sorted by descending sigma like JAMA. The
// n/2 disjoint pairs of one round of a round-robin schedule are independent: one CTA per pair, n - 1 launches per
// sweep, sweeps until a whole sweep rotates nothing. U, S, V agree with the reference's up to the sign of each
// singular-vector pair (and the order within repeated singular values); U S V^T = A either way.

__global__ void __launch_bounds__(256) svd_load_kernel(const double *__restrict__ src, int64_t rs, int64_t cs, double *w,
        double *v, int m, int n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x, tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < (int64_t) m * n; i += stride) {
        const int64_t col = i / m, row = i - col * m;
        w[i] = src[row * rs + col * cs];
    }
    for (int64_t i = tid; i < (int64_t) n * n; i += stride) {
        v[i] = (i / n == i % n) ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ double block_sum_256(double x, double *sh) {
    for (int off = 16; off > 0; off >>= 1) {
        x += __shfl_down_sync(0xffffffffu, x, off);
    }
    __syncthreads();  // sh may still be read from a previous call
    if ((threadIdx.x & 31) == 0) {
        sh[threadIdx.x >> 5] = x;
    }
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; w++) {
        t += sh[w];
    }
    return t;
}

// Round `r` of the circle schedule over n2 (even) players: pair k is (n2-1, r) for k = 0, else
// ((r + k) mod (n2-1), (r - k) mod (n2-1)); players >= n are byes.
__global__ void __launch_bounds__(256) svd_round_kernel(double *w, double *v, int m, int n, int n2, int r, double tol,
        int *rotations) {
    __shared__ double sh[8];
    __shared__ double cs[2];
    const int k = blockIdx.x;
    const int a = k == 0 ? n2 - 1 : (r + k) % (n2 - 1), b = k == 0 ? r : (r - k + n2 - 1) % (n2 - 1);
    const int p = a < b ? a : b, q = a < b ? b : a;
    if (q >= n) {
        return;
    }
    double *wp = w + (int64_t) p * m, *wq = w + (int64_t) q * m;
    double al = 0.0, be = 0.0, ga = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        const double x = wp[i], y = wq[i];
        al += x * x;
        be += y * y;
        ga += x * y;
    }
    al = block_sum_256(al, sh);
    be = block_sum_256(be, sh);
    ga = block_sum_256(ga, sh);
    if (!(fabs(ga) > tol * sqrt(al * be)) || ga == 0.0) {
        return;  // uniform: every thread holds the same sums
    }
    if (threadIdx.x == 0) {
        const double zeta = (be - al) / (2.0 * ga);
        const double t = zeta == 0.0 ? 1.0 : copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        cs[0] = c;
        cs[1] = c * t;
        atomicAdd(rotations, 1);
    }
    __syncthreads();
    const double c = cs[0], sn = cs[1];
    for (int i = threadIdx.x; i < m; i += 256) {
        const double x = wp[i], y = wq[i];
        wp[i] = c * x - sn * y;
        wq[i] = sn * x + c * y;
    }
    double *vp = v + (int64_t) p * n, *vq = v + (int64_t) q * n;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double x = vp[i], y = vq[i];
        vp[i] = c * x - sn * y;
        vq[i] = sn * x + c * y;
    }
}

__global__ void __launch_bounds__(256) svd_norm_kernel(const double *__restrict__ w, double *sigma, int m) {
    __shared__ double sh[8];
    const double *col = w + (int64_t) blockIdx.x * m;
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) {
        acc += col[i] * col[i];
    }
    acc = block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        sigma[blockIdx.x] = sqrt(acc);
    }
}

// Row-major outputs in sorted order: U (m x n), S (n), V (n x n).
__global__ void __launch_bounds__(256) svd_store_kernel(const double *__restrict__ w, const double *__restrict__ v,
        const double *__restrict__ sigma, const int32_t *__restrict__ perm, double *u_out, double *s_out, double *v_out,
        int m, int n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x, tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = tid; i < (int64_t) m * n; i += stride) {
        const int64_t row = i / n, jj = i - row * n;
        const int j = perm[jj];
        const double sg = sigma[j];
        u_out[i] = sg > 0.0 ? w[(int64_t) j * m + row] / sg : 0.0;
    }
    for (int64_t i = tid; i < (int64_t) n * n; i += stride) {
        const int64_t row = i / n, jj = i - row * n;
        v_out[i] = v[(int64_t) perm[jj] * n + row];
    }
    for (int64_t i = tid; i < n; i += stride) {
        s_out[i] = sigma[perm[i]];
    }
}

static thread_local Scratch t_svd_w, t_svd_v, t_svd_s, t_svd_perm;

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

int sstc_svd_dev(const double *src, int64_t nsrc, int src_stride_row, int src_stride_col, double *u, int64_t nu, double *sv,
        int64_t ns, double *v, int64_t nv, int n_rows, int n_cols, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        const int m = n_rows, n = n_cols;
        // LinearAlgebraOpsSvd.cpp:42-50: tall or square, row-major or its transpose
        if (m < 0 || n < 0 || m < n || nsrc != (int64_t) m * n || nu != (int64_t) m * n || ns != n || nv != (int64_t) n * n ||
                !((src_stride_row == n && src_stride_col == 1) || (src_stride_row == 1 && src_stride_col == m)) ||
                (nsrc > 0 && (!src || !u)) || (n > 0 && (!sv || !v))) {
            throw std::runtime_error("Invalid arguments");
        }
        if (n == 0) {
            return;
        }
        double *w = static_cast<double *>(t_svd_w.get((int64_t) m * n * 8));
        double *vv = static_cast<double *>(t_svd_v.get((int64_t) n * n * 8));
        double *sigma = static_cast<double *>(t_svd_s.get((int64_t) n * 8));
        int32_t *perm = static_cast<int32_t *>(t_svd_perm.get((int64_t) n * 4));
        int *rot = static_cast<int *>(t_small.get(sizeof(int)));
        const int grid = stream_grid((int64_t) m * n);
        svd_load_kernel<<<grid, 256, 0, s>>>(src, src_stride_row, src_stride_col, w, vv, m, n);
        check_launch("svd_load_kernel");
        const int n2 = n + (n & 1);
        for (int sweep = 0; sweep < 60 && n > 1; sweep++) {
            SSTC_CUDA(cudaMemsetAsync(rot, 0, sizeof(int), s));
            for (int r = 0; r < n2 - 1; r++) {
                svd_round_kernel<<<n2 / 2, 256, 0, s>>>(w, vv, m, n, n2, r, 1e-15, rot);
            }
            check_launch("svd_round_kernel");
            count_launch(n2 - 2);
            int rotated = 0;
            SSTC_CUDA(cudaMemcpyAsync(&rotated, rot, sizeof(int), cudaMemcpyDeviceToHost, s));
            SSTC_CUDA(cudaStreamSynchronize(s));
            if (rotated == 0) {
                break;
            }
        }
        svd_norm_kernel<<<n, 256, 0, s>>>(w, sigma, m);
        check_launch("svd_norm_kernel");
        std::vector<double> hs((size_t) n);
        SSTC_CUDA(cudaMemcpyAsync(hs.data(), sigma, (size_t) n * 8, cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
        std::vector<int32_t> hp((size_t) n);
        for (int i = 0; i < n; i++) {
            hp[(size_t) i] = i;
        }
        std::stable_sort(hp.begin(), hp.end(), [&](int32_t a, int32_t b) { return hs[(size_t) a] > hs[(size_t) b]; });
        SSTC_CUDA(cudaMemcpyAsync(perm, hp.data(), (size_t) n * 4, cudaMemcpyHostToDevice, s));
        svd_store_kernel<<<grid, 256, 0, s>>>(w, vv, sigma, perm, u, sv, v, m, n);
        check_launch("svd_store_kernel");
        SSTC_CUDA(cudaStreamSynchronize(s));  // hp goes out of scope
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

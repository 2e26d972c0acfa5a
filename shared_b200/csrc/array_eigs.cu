// array_eigs.cu -- ArrayKernel.eigs: eigenvectors and eigenvalues of a general
// real matrix, replacing native/src/shared/LinearAlgebraOpsEigs.cpp:26-615 (Bindings.cpp:134-137).
//
// The arithmetic lives in eigs_core.cuh (shared with the host emulation the CPU tests run under ThreadSanitizer);
// this file holds the three launches and the two entry points:
//   1. eigs_schur_kernel     ONE CTA: Hessenberg reduction + accumulation, then the double-shift QR on H alone (its
//                            steps are recorded), then the recorded steps applied to V. The working matrix sits in
//                            shared memory (row stride padded to an odd count, so walking a row and walking a column
//                            are both bank-conflict free) when size <= 158, and V takes its place there for the last
//                            stage; larger matrices work in their HBM buffers.
//   2. eigs_backsub_kernel   one thread per eigenvalue: back-substitution against the read-only Schur form.
//   3. eigs_backtransform_kernel  one thread per element of vec = V X.
// Built with -fmad=false (Makefile NOFMA): with every element updated in the reference's operation order the output is
// bit-identical to the reference's. Like invert and svd this is a correctness-first kernel set off the hot path.
#include <algorithm>
#include <string>

#include "array_common.cuh"
#include "eigs_core.cuh"

namespace sstc {

struct DevCta {
    int tid, nthreads;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void atomic_max(int *p, int v) { atomicMax(p, v); }
};

constexpr int kEigsMaxThreads = 512;
constexpr int64_t kEigsSmemBudget = 200 * 1024;  // dynamic; + 12 KB of static staging, of the 227 KB a CTA may opt into
constexpr int kEigsChunk = 256;                  // recorded steps staged per pass over V

__global__ void __launch_bounds__(kEigsMaxThreads) eigs_schur_kernel(const double *__restrict__ src, double *T,
        double *V, double *val, double *ort, double *norm, int *status, EigsRot *log, int log_cap, int n, int in_smem) {
    extern __shared__ double s_h[];
    __shared__ EigsShared sh;
    __shared__ EigsRot chunk[kEigsChunk];
    DevCta cta{(int) threadIdx.x, (int) blockDim.x};
    double *H = in_smem ? s_h : T;
    const int ld = in_smem ? (n | 1) : n;
    for (int e = cta.tid; e < n * n; e += cta.nthreads) {
        H[(e / n) * ld + e % n] = src[e];
    }
    cta.sync();
    eigs_hessenberg(cta, H, ld, V, ort, n, &sh);
    cta.sync();
    const int pending = eigs_schur(cta, H, ld, V, val, n, &sh, norm, log, log_cap, chunk, kEigsChunk);  // uniform
    cta.sync();
    if (cta.tid == 0) {
        *status = pending < 0;
    }
    if (in_smem) {
        for (int e = cta.tid; e < n * n; e += cta.nthreads) {
            T[e] = H[(e / n) * ld + e % n];
        }
        cta.sync();
        if (pending > 0) {  // V takes H's place in shared memory, the recorded steps are applied, V goes back
            for (int e = cta.tid; e < n * n; e += cta.nthreads) {
                s_h[(e / n) * ld + e % n] = V[e];
            }
            cta.sync();
            eigs_apply_rots(cta, s_h, ld, n, log, pending, chunk, kEigsChunk);
            for (int e = cta.tid; e < n * n; e += cta.nthreads) {
                V[e] = s_h[(e / n) * ld + e % n];
            }
        }
    } else if (pending > 0) {
        eigs_apply_rots(cta, V, n, n, log, pending, chunk, kEigsChunk);
    }
}

__global__ void __launch_bounds__(32) eigs_backsub_kernel(const double *__restrict__ T, double *X,
        const double *__restrict__ val, const double *__restrict__ norm, const int *__restrict__ status, int n) {
    const double nrm = *norm;
    const int c = blockIdx.x * 32 + threadIdx.x;
    if (*status != 0 || nrm == 0.0 || c >= n) {  // LinearAlgebraOpsEigs.cpp:457-459
        return;
    }
    eigs_backsub_column(T, X, val, n, c, nrm);
}

__global__ void __launch_bounds__(kThreads) eigs_backtransform_kernel(const double *__restrict__ V,
        const double *__restrict__ X, const double *__restrict__ norm, const int *__restrict__ status, double *vec,
        int n) {
    if (*status != 0) {
        return;
    }
    const double nrm = *norm;
    const int total = n * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        vec[e] = nrm == 0.0 ? V[e] : eigs_backtransform_element(V, X, n, e / n, e % n);
    }
}

static thread_local Scratch t_eigs_t, t_eigs_v, t_eigs_x, t_eigs_small, t_eigs_log;

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_eigs_dev(const double *src, int64_t nsrc, double *vec, int64_t nvec, double *val, int64_t nval, int size,
        void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        const int64_t n2 = (int64_t) size * size;
        // LinearAlgebraOpsEigs.cpp:33-43; n2 has to fit a Java array
        if (size < 0 || n2 > INT32_MAX || nsrc != n2 || nvec != n2 || nval != 2 * (int64_t) size ||
                (n2 > 0 && (!src || !vec || !val))) {
            throw std::runtime_error("Invalid arguments");
        }
        if (size == 0) {
            return;
        }
        const int n = size;
        double *T = static_cast<double *>(t_eigs_t.get(n2 * 8));
        double *V = static_cast<double *>(t_eigs_v.get(n2 * 8));
        double *X = static_cast<double *>(t_eigs_x.get(n2 * 8));
        double *small = static_cast<double *>(t_eigs_small.get(((int64_t) n + 3) * 8));
        double *norm = small, *ort = small + 2;
        int *status = reinterpret_cast<int *>(small + 1);
        const int64_t smem = (int64_t) n * (n | 1) * 8;
        const int in_smem = smem <= kEigsSmemBudget;
        if (in_smem) {
            SSTC_CUDA(cudaFuncSetAttribute(eigs_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                    (int) kEigsSmemBudget));
        }
        // room for the recorded QR steps (about 1.5 n^2 of them on random input); a full log is applied on the way
        const int log_cap = (int) std::min<int64_t>(std::max<int64_t>(8 * n2, 4096), 1 << 21);
        EigsRot *log = static_cast<EigsRot *>(t_eigs_log.get((int64_t) log_cap * (int64_t) sizeof(EigsRot)));
        const int threads = std::min(kEigsMaxThreads, std::max(32, (n + 31) / 32 * 32));
        SSTC_CUDA(cudaMemsetAsync(X, 0, (size_t) n2 * 8, s));
        eigs_schur_kernel<<<1, threads, in_smem ? (size_t) smem : 0, s>>>(src, T, V, val, ort, norm, status, log, log_cap,
                n, in_smem);
        check_launch("eigs_schur_kernel");
        eigs_backsub_kernel<<<(n + 31) / 32, 32, 0, s>>>(T, X, val, norm, status, n);
        check_launch("eigs_backsub_kernel");
        eigs_backtransform_kernel<<<stream_grid(n2), kThreads, 0, s>>>(V, X, norm, status, vec, n);
        check_launch("eigs_backtransform_kernel");
        int failed = 0;
        SSTC_CUDA(cudaMemcpyAsync(&failed, status, sizeof(int), cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
        if (failed) {
            throw std::runtime_error("Eigenvalue iteration did not converge");
        }
    });
}

/* Host tier: Bindings.cpp:134-137 -> LinearAlgebraOps::eigs. srcV is pinned read-only there (no copy-back). */
int sstc_eigs(const double *src, int64_t nsrc, double *vec, int64_t nvec, double *val, int64_t nval, int size) {
    return guarded([&] {
        if (nsrc < 0 || nvec < 0 || nval < 0 || (nsrc > 0 && !src) || (nvec > 0 && !vec) || (nval > 0 && !val)) {
            throw std::runtime_error("Invalid arguments");
        }
        static thread_local Scratch h_src, h_vec, h_val;
        double *ds = static_cast<double *>(h_src.get(std::max<int64_t>(nsrc, 1) * 8));
        double *dv = static_cast<double *>(h_vec.get(std::max<int64_t>(nvec, 1) * 8));
        double *dl = static_cast<double *>(h_val.get(std::max<int64_t>(nval, 1) * 8));
        if (nsrc > 0) {
            SSTC_CUDA(cudaMemcpyAsync(ds, src, (size_t) nsrc * 8, cudaMemcpyHostToDevice, 0));
        }
        if (sstc_eigs_dev(ds, nsrc, dv, nvec, dl, nval, size, nullptr) != 0) {
            throw std::runtime_error(std::string(sstc_last_error()));
        }
        if (nvec > 0 && size > 0) {
            SSTC_CUDA(cudaMemcpyAsync(vec, dv, (size_t) nvec * 8, cudaMemcpyDeviceToHost, 0));
            SSTC_CUDA(cudaMemcpyAsync(val, dl, (size_t) nval * 8, cudaMemcpyDeviceToHost, 0));
        }
        SSTC_CUDA(cudaStreamSynchronize(0));
    });
}

}  // extern "C"

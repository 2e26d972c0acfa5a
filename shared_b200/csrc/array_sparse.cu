// array_sparse.cu -- ArrayKernel.insertSparse / sliceSparse on the GPU.
//
// The algorithm, its per-element code and its orchestration live in sparse_core.cuh (shared with the serial host
// executor of tests/emu/sparse_emu.cpp, which checks them against the reference's native SparseOps on the CPU). This
// file supplies the CUDA executor -- grid-stride kernels over the functors, the three-phase int32 -> int64 exclusive
// scan, the bitonic (key, position) pair sort of array_dimension.cu, per-thread grow-only scratch -- and the two host-
// tier entry points. All bulk work (sorting, merging, metadata, value movement) runs on the device; the host builds
// only the per-call slice lookup tables, whose size is that of the slice specification.
#include <string>

#include "array_common.cuh"
#include "sparse_core.cuh"

namespace sstc {

struct SpDevCta {
    int tid, nthreads;
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <class F>
__global__ void __launch_bounds__(kThreads) sp_for_each_kernel(int64_t n, F f) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        f(i);
    }
}

__global__ void __launch_bounds__(kScanThreads) sp_scan_sum_kernel(const int32_t *__restrict__ in, int64_t n,
        int64_t n_chunks, int64_t *chunk_sums) {
    __shared__ int64_t smem[kScanThreads];
    SpDevCta cta{(int) threadIdx.x, (int) blockDim.x};
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        scan_chunk_sum(cta, in, n, chunk, chunk_sums, smem);
    }
}

__global__ void __launch_bounds__(kScanThreads) sp_scan_offsets_kernel(int64_t *chunk_sums, int64_t n_chunks,
        int64_t *total) {
    __shared__ int64_t smem[kScanThreads];
    SpDevCta cta{(int) threadIdx.x, (int) blockDim.x};
    scan_chunk_offsets(cta, chunk_sums, n_chunks, total, smem);
}

__global__ void __launch_bounds__(kScanThreads) sp_scan_write_kernel(const int32_t *__restrict__ in, int64_t n,
        int64_t n_chunks, const int64_t *__restrict__ chunk_offsets, int64_t *out) {
    __shared__ int64_t smem[kScanThreads];
    SpDevCta cta{(int) threadIdx.x, (int) blockDim.x};
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        scan_chunk_write(cta, in, n, chunk, chunk_offsets, out, smem);
    }
}

static thread_local Scratch t_sp[SPB_SLOTS];
static thread_local Scratch t_sp_sums;

// The executor sparse_core.cuh is written against; everything goes to the legacy default stream like the rest of the
// host tier (array_host.cu).
struct CudaSparseExec {
    cudaStream_t s = 0;

    template <class T>
    T *buf(int slot, int64_t n) {
        return static_cast<T *>(t_sp[slot].get(std::max<int64_t>(n, 1) * (int64_t) sizeof(T)));
    }

    template <class F>
    void for_each(int64_t n, const F &f) {
        if (n <= 0) {
            return;
        }
        sp_for_each_kernel<F><<<stream_grid(n), kThreads, 0, s>>>(n, f);
        check_launch("sp_for_each_kernel");
    }

    void scan(const int32_t *in, int64_t n, int64_t *out) {
        if (n <= 0) {
            SSTC_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
            return;
        }
        const int64_t n_chunks = (n + kScanChunk - 1) / kScanChunk;
        int64_t *sums = static_cast<int64_t *>(t_sp_sums.get(n_chunks * 8));
        const int grid = (int) std::min<int64_t>(n_chunks, (int64_t) sm_count() * kCtasPerSm);
        sp_scan_sum_kernel<<<grid, kScanThreads, 0, s>>>(in, n, n_chunks, sums);
        check_launch("sp_scan_sum_kernel");
        sp_scan_offsets_kernel<<<1, kScanThreads, 0, s>>>(sums, n_chunks, out + n);
        check_launch("sp_scan_offsets_kernel");
        sp_scan_write_kernel<<<grid, kScanThreads, 0, s>>>(in, n, n_chunks, sums, out);
        check_launch("sp_scan_write_kernel");
    }

    void sort_pairs(uint64_t *keys, int32_t *pay, int64_t total, int64_t seg) {
        sort_pairs_segments(keys, pay, total, seg, s);
    }

    template <class T>
    T read(const T *p) {
        T v;
        SSTC_CUDA(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
        return v;
    }

    void zero(void *p, int64_t bytes) {
        if (bytes > 0) {
            SSTC_CUDA(cudaMemsetAsync(p, 0, (size_t) bytes, s));
        }
    }

    void upload(void *dst, const void *src, int64_t bytes) {
        if (bytes > 0) {
            SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyHostToDevice, s));
        }
    }

    void download(void *dst, const void *src, int64_t bytes) {
        if (bytes > 0) {
            SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyDeviceToHost, s));
        }
    }

    void sync() { SSTC_CUDA(cudaStreamSynchronize(s)); }
};

// the arrays a call hands back stay valid until the calling thread's next sparse call
static thread_local SparseHostResult t_result;

static void publish(const SparseHostResult &r, int elem_bytes, sstc_sparse_state *out) {
    out->count = r.count;
    out->n_offsets = (int32_t) r.offsets.size();
    out->n_indirections = (int32_t) r.indirections.size();
    out->elem_bytes = elem_bytes;
    out->values = r.values.data();
    out->indices = r.indices.data();
    out->indirection_offsets = r.offsets.data();
    out->indirections = r.indirections.data();
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_insert_sparse(int elem_bytes, const void *oldV, int64_t nold, const int32_t *oldD, const int32_t *oldS,
        const int32_t *oldDo, int nd, const int32_t *oldI, const void *newV, int64_t nnew, int32_t *newI,
        sstc_sparse_state *out) {
    return guarded([&] {
        if (!out) {
            throw std::runtime_error("Invalid arguments");
        }
        memset(out, 0, sizeof(*out));
        CudaSparseExec ex;
        sp_insert(ex, elem_bytes, oldV, nold, oldD, oldS, oldDo, nd, oldI, newV, nnew, newI, t_result);
        publish(t_result, elem_bytes, out);
    });
}

int sstc_slice_sparse(int elem_bytes, const int32_t *slices, int nslices3, const void *srcV, int64_t nsrc,
        const int32_t *srcD, const int32_t *srcS, const int32_t *srcDo, const int32_t *srcI, const int32_t *srcIo,
        const int32_t *srcIi, const void *dstV, int64_t ndst, const int32_t *dstD, const int32_t *dstS,
        const int32_t *dstDo, const int32_t *dstI, const int32_t *dstIo, const int32_t *dstIi, int nd,
        sstc_sparse_state *out) {
    return guarded([&] {
        if (!out) {
            throw std::runtime_error("Invalid arguments");
        }
        memset(out, 0, sizeof(*out));
        CudaSparseExec ex;
        sp_slice(ex, elem_bytes, slices, nslices3, srcV, nsrc, srcD, srcS, srcDo, srcI, srcIo, srcIi, dstV, ndst, dstD,
                dstS, dstDo, dstI, dstIo, dstIi, nd, t_result);
        publish(t_result, elem_bytes, out);
    });
}

}  // extern "C"

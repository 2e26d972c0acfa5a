// array_image.cu -- ImageKernel half of the C ABI: integral images and integral histograms.
//
// Replaces src/org/shared/image/kernel/ImageOps.java:71-199 (JavaImageKernel) and
// native/src/shared/NativeImageKernel.cpp:32-247 (NativeImageKernel, JNI exports Bindings.cpp:144-158), the
// second native service the reference's library registers (Library.cpp:64-66). Semantics kept exactly:
//   * dst has every source dimension + 1 (histogram: plus a trailing "bins" dimension); the source is copied to
//     offset +1 along every dimension, nothing else of dst is cleared (callers pass zero-filled arrays; whatever
//     the border and the other bins hold takes part in the sums, as in the reference);
//   * then an inclusive running sum along dimension 0, 1, ... in turn over the WHOLE dst (all bins).
// Here: one modular block copy (sstc_map_dev) or a membership scatter kernel, then the rdOp scan kernels in
// place -- no index tables (the reference builds two int[] tables of srcLen + dstLen entries per call).
#include <algorithm>
#include <string>
#include <vector>

#include "array_common.cuh"

namespace sstc {

struct ScatterSpec {
    int nd;
    int64_t total;
    int64_t size[kMaxDims];  // source dims, fastest source stride first
    int64_t sstride[kMaxDims];
    int64_t dstride[kMaxDims];
    int64_t dst_offset, bin_stride;
    int32_t nbins;
};

// Thread per source element: validates the membership (first bad one sets *bad) or scatters the value to its bin.
template <bool CHECK>
__global__ void __launch_bounds__(kThreads) histogram_scatter_kernel(const double *__restrict__ src,
        const int32_t *__restrict__ mem, double *dst, ScatterSpec S, int *bad) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < S.total; i += stride) {
        int64_t rem = i, soff = 0, doff = S.dst_offset;
        for (int k = 0; k < S.nd; k++) {
            const int64_t q = rem / S.size[k];
            const int64_t r = rem - q * S.size[k];
            soff += r * S.sstride[k];
            doff += r * S.dstride[k];
            rem = q;
        }
        const int32_t bin = mem[soff];
        if (CHECK) {
            if (!(bin >= 0 && bin < S.nbins)) {
                *bad = 1;
            }
        } else {
            dst[doff + (int64_t) bin * S.bin_stride] = src[soff];
        }
    }
}

static void rethrow_rc(int rc) {
    if (rc != 0) {
        throw std::runtime_error(std::string(sstc_last_error()));
    }
}

static void check_dims(const int32_t *dims, const int32_t *strides, int nd, int64_t len) {
    // MappingOps::checkDimensions (MappingOps.cpp:73-91)
    int64_t acc = 0;
    for (int i = 0; i < nd; i++) {
        if (dims[i] < 0 || strides[i] < 0) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
        acc += (int64_t) (dims[i] - 1) * strides[i];
    }
    if (acc != len - 1) {
        throw std::runtime_error("Invalid dimensions and/or strides");
    }
}

static thread_local Scratch t_flag;

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_integral_image_dev(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, void *stream) {
    return guarded([&] {
        if (!sD || !sS || !dD || !dS || nd < 0 || nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nd > kMaxDims) {
            throw std::runtime_error("Too many dimensions for the CUDA provider");
        }
        check_dims(sD, sS, nd, nsrc);
        check_dims(dD, dS, nd, ndst);
        for (int d = 0; d < nd; d++) {
            if (sD[d] + 1 != dD[d]) {
                throw std::runtime_error("Dimension mismatch");  // NativeImageKernel.cpp:73-76
            }
        }
        if (nsrc == 0) {
            return;
        }
        // the source block lands at (1, 1, ...): map bounds are (srcStart, dstStart, size) per dimension
        std::vector<int32_t> bounds((size_t) 3 * nd), op((size_t) nd);
        for (int d = 0; d < nd; d++) {
            bounds[(size_t) 3 * d] = 0;
            bounds[(size_t) 3 * d + 1] = 1;
            bounds[(size_t) 3 * d + 2] = sD[d];
            op[(size_t) d] = d;
        }
        rethrow_rc(sstc_map_dev(8, bounds.data(), 3 * nd, src, nsrc, sD, sS, dst, ndst, dD, dS, nd, stream));
        rethrow_rc(sstc_rdop_dev(RD_SUM, dst, ndst, dD, dS, nd, dst, ndst, op.data(), nd, stream));
    });
}

int sstc_integral_histogram_dev(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd,
        const int32_t *mem, int64_t nmem, double *dst, int64_t ndst, const int32_t *dD, const int32_t *dS,
        void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        if (!sD || !sS || !dD || !dS || nd < 0 || nsrc < 0 || ndst < 0 || nsrc != nmem || (nsrc > 0 && (!src || !mem)) ||
                (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nd + 1 > kMaxDims) {
            throw std::runtime_error("Too many dimensions for the CUDA provider");
        }
        check_dims(sD, sS, nd, nsrc);
        check_dims(dD, dS, nd + 1, ndst);
        for (int d = 0; d < nd; d++) {
            if (sD[d] + 1 != dD[d]) {
                throw std::runtime_error("Dimension mismatch");
            }
        }
        if (nsrc == 0) {
            return;
        }
        ScatterSpec S;
        S.nd = 0;
        S.total = 1;
        S.dst_offset = 0;
        std::vector<int> order;
        for (int d = 0; d < nd; d++) {
            S.dst_offset += dS[d];
            S.total *= sD[d];
            if (sD[d] > 1) {
                order.push_back(d);
            }
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sS[a] < sS[b]; });
        for (int d : order) {
            S.size[S.nd] = sD[d];
            S.sstride[S.nd] = sS[d];
            S.dstride[S.nd] = dS[d];
            S.nd++;
        }
        S.nbins = dD[nd];
        S.bin_stride = dS[nd];
        // The reference throws on the first bad membership (NativeImageKernel.cpp:190-194); here all of them are
        // checked before dst is touched.
        int *bad = static_cast<int *>(t_flag.get(sizeof(int)));
        SSTC_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
        const int grid = stream_grid(S.total);
        histogram_scatter_kernel<true><<<grid, kThreads, 0, s>>>(src, mem, dst, S, bad);
        check_launch("histogram_scatter_kernel");
        int bad_h = 0;
        SSTC_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        SSTC_CUDA(cudaStreamSynchronize(s));
        if (bad_h) {
            throw std::runtime_error("Invalid membership index");
        }
        histogram_scatter_kernel<false><<<grid, kThreads, 0, s>>>(src, mem, dst, S, bad);
        check_launch("histogram_scatter_kernel");
        std::vector<int32_t> op((size_t) nd);
        for (int d = 0; d < nd; d++) {
            op[(size_t) d] = d;
        }
        rethrow_rc(sstc_rdop_dev(RD_SUM, dst, ndst, dD, dS, nd + 1, dst, ndst, op.data(), nd, stream));
    });
}

}  // extern "C"

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

// First scan of the integral image fused with the copy: the running sum along dimension 0 of "dst with the source
// block placed at +1", without materialising that array first. A CTA owns a strip of 32 adjacent destination lines
// (lane = line) and walks down them in blocks of 8 warps x kImgRows rows, like the rdOp strip scan
// (array_dimension.cu). Row 0 of a line is the destination's own border value; rows >= 1 come from the source when
// the line is an interior one (every other coordinate >= 1) and from the destination otherwise -- so whatever the
// border holds takes part exactly as in the reference (NativeImageKernel.cpp:96-118).
constexpr int kImgRows = 32;

struct ImgLines {
    int nd;                      // the other dimensions (destination sizes), fastest destination stride first
    int64_t size[kMaxDims], dstride[kMaxDims], sstride[kMaxDims];
    int64_t n_lines, len;        // len = destination size along dimension 0
    int64_t dlstride, slstride;  // strides along dimension 0
};

__global__ void __launch_bounds__(kThreads) integral_first_scan_kernel(const double *__restrict__ src, double *dst,
        ImgLines L) {
    constexpr int NW = kThreads / 32;
    __shared__ double tot[2][NW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t strips = (L.n_lines + 31) / 32;
    for (int64_t sidx = blockIdx.x; sidx < strips; sidx += gridDim.x) {
        int64_t m = sidx * 32 + lane;
        const bool live = m < L.n_lines;
        int64_t doff = 0, soff = 0;
        bool interior = live;
        if (live) {
            for (int k = 0; k < L.nd; k++) {
                const int64_t q = m / L.size[k];
                const int64_t r = m - q * L.size[k];
                doff += r * L.dstride[k];
                soff += (r - 1) * L.sstride[k];
                interior = interior && r >= 1;
                m = q;
            }
        }
        double carry = 0.0;
        int buf = 0;
        for (int64_t base = 0; base < L.len; base += (int64_t) NW * kImgRows, buf ^= 1) {
            const int64_t r0 = base + (int64_t) warp * kImgRows;
            double x[kImgRows];
#pragma unroll
            for (int e = 0; e < kImgRows; e++) {
                const int64_t r = r0 + e;
                double v = 0.0;
                if (live && r < L.len) {
                    v = (interior && r >= 1) ? src[soff + (r - 1) * L.slstride] : dst[doff + r * L.dlstride];
                }
                x[e] = v;
            }
#pragma unroll
            for (int e = 1; e < kImgRows; e++) {
                x[e] = x[e - 1] + x[e];
            }
            tot[buf][warp][lane] = x[kImgRows - 1];
            __syncthreads();
            double pre = carry;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const double t = tot[buf][w][lane];
                pre = w < warp ? pre + t : pre;
                carry = carry + t;
            }
#pragma unroll
            for (int e = 0; e < kImgRows; e++) {
                if (live && r0 + e < L.len) {
                    dst[doff + (r0 + e) * L.dlstride] = pre + x[e];
                }
            }
        }
        __syncthreads();
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
        // 2-D dense images: ONE pass -- every source element read once, every destination element written once; the
        // copy to offset (1, 1) and both scans happen in the tile (array_scan2d.cu)
        if (nd == 2 && sD[0] > 0 && sD[1] > 0) {
            const bool row_major = dS[1] == 1 && dS[0] == dD[1] && sS[1] == 1 && sS[0] == sD[1];
            const bool col_major = dS[0] == 1 && dS[1] == dD[0] && sS[0] == 1 && sS[1] == sD[0];
            const int64_t rows = row_major ? dD[0] : dD[1], cols = row_major ? dD[1] : dD[0];
            if ((row_major || col_major) && scan2d_applicable(rows, cols)) {
                scan2d_sum(src, cols - 1, dst, cols, rows, cols, true, as_stream(stream));
                return;
            }
        }
        // Fused path: dimension 0 is not the fastest one and there are enough lines for coalesced strips -- the copy
        // is folded into the first scan (8 B read + 8 B written per point instead of 32), the other dimensions are
        // scanned in place afterwards.
        int fastest = 0;
        int64_t n_lines = 1;
        for (int d = 1; d < nd; d++) {
            n_lines *= dD[d];
            if (dS[d] < dS[fastest]) {
                fastest = d;
            }
        }
        if (nd >= 2 && fastest != 0 && n_lines >= 1024 && dD[0] >= 2) {
            ImgLines L;
            L.nd = 0;
            L.n_lines = n_lines;
            L.len = dD[0];
            L.dlstride = dS[0];
            L.slstride = sS[0];
            std::vector<int> order;
            for (int d = 1; d < nd; d++) {
                order.push_back(d);
            }
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dS[a] < dS[b]; });
            for (int d : order) {
                L.size[L.nd] = dD[d];
                L.dstride[L.nd] = dS[d];
                L.sstride[L.nd] = sS[d];
                L.nd++;
            }
            const int grid = (int) std::min<int64_t>((n_lines + 31) / 32, (int64_t) sm_count() * 8);
            integral_first_scan_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(src, dst, L);
            check_launch("integral_first_scan_kernel");
            std::vector<int32_t> op;
            for (int d = 1; d < nd; d++) {
                op.push_back(d);
            }
            rethrow_rc(sstc_rdop_dev(RD_SUM, dst, ndst, dD, dS, nd, dst, ndst, op.data(), (int) op.size(), stream));
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

// array_structural.cu -- the structural callers of ArrayKernel.map / slice as single device-tier entry points.
//
// In the reference these live on the JVM side: each computes bounds or slice triples and makes ONE kernel call --
//   ProtoArray.shift      (src/org/shared/array/ProtoArray.java:313-337)   map with bounds (0, shift, size) per dimension
//   ProtoArray.transpose  (ProtoArray.java:246-283)                         map onto the destination's permuted strides
//   AbstractComplexArray.fftShift / ifftShift (AbstractComplexArray.java:67-97)   shift by +-dims/2, complex pairs kept
//   ConvolutionCache.pad  (src/org/shared/fft/ConvolutionCache.java:149-239)     mirror padding, 2^nDims splice calls
// A device-resident caller (shared_b200/device_array.py; a CudaRealArray on the Java side) wants them without building
// int[] specifications per call, so they are exported here; they run on the map / slice kernels of array_mapping.cu
// (row-walking copy, 32 x 32 tile transpose), measured at 0.8-0.97 of HBM at 8192 x 8192 (bench.py, C4).
#include <vector>

#include "array_common.cuh"

namespace sstc {

static void rethrow(int rc) {
    if (rc != 0) {
        throw std::runtime_error(std::string(sstc_last_error()));
    }
}

static void check_nd(const int32_t *dims, int nd) {
    if (!dims || nd <= 0 || nd > kMaxDims) {
        throw std::runtime_error("Invalid arguments");
    }
    for (int d = 0; d < nd; d++) {
        if (dims[d] < 0) {
            throw std::runtime_error("Invalid dimensions and/or strides");
        }
    }
}

static std::vector<int32_t> strides_for(const int32_t *dims, int nd, int near_order) {
    // IndexingOrder.strides (Array.java:240-287): FAR = row-major, NEAR = column-major
    std::vector<int32_t> s((size_t) nd, 1);
    if (!near_order) {
        for (int i = nd - 2; i >= 0; i--) {
            s[(size_t) i] = s[(size_t) i + 1] * dims[i + 1];
        }
    } else {
        for (int i = 1; i < nd; i++) {
            s[(size_t) i] = s[(size_t) i - 1] * dims[i - 1];
        }
    }
    return s;
}

static int64_t count(const int32_t *dims, int nd) {
    int64_t n = 1;
    for (int d = 0; d < nd; d++) {
        n *= dims[d];
    }
    return n;
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_shift_dev(int elem_bytes, const void *src, void *dst, const int32_t *dims, int nd, int near_order, const int32_t *shifts,
        void *stream) {
    return guarded([&] {
        check_nd(dims, nd);
        if (!shifts) {
            throw std::runtime_error("Invalid arguments");
        }
        const std::vector<int32_t> st = strides_for(dims, nd, near_order);
        std::vector<int32_t> bounds((size_t) 3 * nd);
        for (int d = 0; d < nd; d++) {
            bounds[(size_t) 3 * d] = 0;
            bounds[(size_t) 3 * d + 1] = shifts[d];
            bounds[(size_t) 3 * d + 2] = dims[d];
        }
        const int64_t n = count(dims, nd);
        rethrow(sstc_map_dev(elem_bytes, bounds.data(), 3 * nd, src, n, dims, st.data(), dst, n, dims, st.data(), nd, stream));
    });
}

int sstc_fftshift_dev(const double *src, double *dst, const int32_t *dims, int nd, int inverse, void *stream) {
    return guarded([&] {
        check_nd(dims, nd);
        if (dims[nd - 1] != 2) {
            throw std::runtime_error("Invalid arguments");  // a ComplexArray's trailing dimension
        }
        std::vector<int32_t> shifts((size_t) nd, 0);
        for (int d = 0; d + 1 < nd; d++) {
            shifts[(size_t) d] = (inverse ? -1 : 1) * (dims[d] / 2);  // AbstractComplexArray.java:88-92
        }
        rethrow(sstc_shift_dev(8, src, dst, dims, nd, 0, shifts.data(), stream));
    });
}

int sstc_transpose_dev(int elem_bytes, const void *src, void *dst, const int32_t *dims, int nd, int near_order,
        const int32_t *permutation, void *stream) {
    return guarded([&] {
        check_nd(dims, nd);
        if (!permutation) {
            throw std::runtime_error("Invalid arguments");
        }
        std::vector<char> seen((size_t) nd, 0);
        std::vector<int32_t> new_dims((size_t) nd);
        for (int d = 0; d < nd; d++) {
            const int p = permutation[d];
            if (p < 0 || p >= nd || seen[(size_t) p]) {
                throw std::runtime_error("Invalid permutation");  // ProtoArray.java:257-262
            }
            seen[(size_t) p] = 1;
            new_dims[(size_t) p] = dims[d];
        }
        const std::vector<int32_t> sst = strides_for(dims, nd, near_order), nst = strides_for(new_dims.data(), nd, near_order);
        std::vector<int32_t> dst_strides((size_t) nd), bounds((size_t) 3 * nd);
        for (int d = 0; d < nd; d++) {
            dst_strides[(size_t) d] = nst[(size_t) permutation[d]];  // the destination seen with the SOURCE's dims
            bounds[(size_t) 3 * d] = 0;
            bounds[(size_t) 3 * d + 1] = 0;
            bounds[(size_t) 3 * d + 2] = dims[d];
        }
        const int64_t n = count(dims, nd);
        rethrow(sstc_map_dev(elem_bytes, bounds.data(), 3 * nd, src, n, dims, sst.data(), dst, n, dims, dst_strides.data(), nd,
                stream));
    });
}

int sstc_pad_dev(const double *src, double *dst, const int32_t *dims, int nd, const int32_t *margins, void *stream) {
    return guarded([&] {
        check_nd(dims, nd);
        if (!margins) {
            throw std::runtime_error("Invalid arguments");
        }
        std::vector<int32_t> out_dims((size_t) nd), slices;
        for (int d = 0; d < nd; d++) {
            const int32_t s = dims[d], m = margins[d];
            if (m < 0 || m > s) {
                throw std::runtime_error("Invalid margin");
            }
            out_dims[(size_t) d] = s + 2 * m;
            for (int32_t j = 0; j < m; j++) {  // left margin, mirrored about the first sample
                slices.insert(slices.end(), {m - 1 - j, j, d});
            }
            for (int32_t k = 0; k < s; k++) {
                slices.insert(slices.end(), {k, m + k, d});
            }
            for (int32_t i = 0; i < m; i++) {  // right margin, mirrored about the last sample
                slices.insert(slices.end(), {s - 1 - i, s + m + i, d});
            }
        }
        const std::vector<int32_t> sst = strides_for(dims, nd, 0), dst_st = strides_for(out_dims.data(), nd, 0);
        rethrow(sstc_slice_dev(8, slices.data(), (int) slices.size(), src, count(dims, nd), dims, sst.data(), dst,
                count(out_dims.data(), nd), out_dims.data(), dst_st.data(), nd, stream));
    });
}

}  // extern "C"

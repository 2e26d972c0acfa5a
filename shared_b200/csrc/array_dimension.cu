// array_dimension.cu -- ArrayKernel dimension operations on the device: reduce (rrOp), index (riOp), prefix
// scan (rdOp), plus the segmented sort behind RI_SORT and the *_SHUFFLE unary ops.
//
// Replaces (paths relative to /root/reference):
//   rrOp   src/org/shared/array/kernel/DimensionOps.java:521-610,72-156   native/src/shared/DimensionOpsReduce.cpp:26-258
//   riOp   DimensionOps.java:615-669,169-378                              native/src/shared/DimensionOpsIndex.cpp:26-350
//   rdOp   DimensionOps.java:674-709,391-485                              native/src/shared/DimensionOpsDimension.cpp:26-207
//
// The reference materialises int[] index tables (MappingOps.assignMappingIndices) and walks one line at a
// time. Here a "line" (all points along the operating dimension for one setting of the other indices) is
// located arithmetically from the dims/strides, and lines are processed in parallel:
//   * operating dimension has the smallest stride  -> one warp or one CTA per line, lanes along the line
//     (coalesced), shuffle / shared-memory fold, scan or ordered compaction;
//   * otherwise                                    -> one thread per line, adjacent threads on adjacent lines
//     (coalesced across lines), sequential along the line exactly as the reference does it.
// Algorithmic HBM traffic: rrOp 8 B/element read (+ output), rdOp 16 B/element per scanned dimension,
// riOp 8 B read (x2 for MAX/MIN) + 4 B written per element.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "array_common.cuh"

namespace sstc {

// ---- line geometry ------------------------------------------------------------------------------------

struct Lines {
    int nd;  // "other" dimensions of size > 1, fastest (smallest input stride) first
    int64_t size[kMaxDims];
    int64_t istride[kMaxDims];
    int64_t ostride[kMaxDims];
    int64_t n_lines, len, ilstride, olstride;
};

__device__ __forceinline__ void line_base(const Lines &L, int64_t m, int64_t &ioff, int64_t &ooff) {
    ioff = 0;
    ooff = 0;
    for (int k = 0; k < L.nd; k++) {
        const int64_t q = m / L.size[k];
        const int64_t r = m - q * L.size[k];
        ioff += r * L.istride[k];
        ooff += r * L.ostride[k];
        m = q;
    }
}

static std::vector<int64_t> far_strides(const std::vector<int64_t> &dims) {
    std::vector<int64_t> s(dims.size(), 1);
    for (int i = (int) dims.size() - 2; i >= 0; i--) {
        s[(size_t) i] = s[(size_t) i + 1] * std::max<int64_t>(dims[(size_t) i + 1], 1);
    }
    return s;
}

// dim < 0: every dimension is an "other" dimension (used by the strided copy)
static Lines make_lines(const std::vector<int64_t> &dims, const std::vector<int64_t> &istr,
        const std::vector<int64_t> &ostr, int dim) {
    Lines L;
    L.nd = 0;
    L.n_lines = 1;
    L.len = dim >= 0 ? dims[(size_t) dim] : 1;
    L.ilstride = dim >= 0 ? istr[(size_t) dim] : 0;
    L.olstride = dim >= 0 ? ostr[(size_t) dim] : 0;
    std::vector<int> order;
    for (int i = 0; i < (int) dims.size(); i++) {
        if (i != dim && dims[(size_t) i] > 1) {
            order.push_back(i);
        }
        if (i != dim) {
            L.n_lines *= dims[(size_t) i];
        }
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return istr[(size_t) a] < istr[(size_t) b]; });
    for (int i : order) {
        L.size[L.nd] = dims[(size_t) i];
        L.istride[L.nd] = istr[(size_t) i];
        L.ostride[L.nd] = ostr[(size_t) i];
        L.nd++;
    }
    return L;
}

static bool line_is_fastest(const Lines &L) {
    return L.len > 1 && (L.nd == 0 || L.ilstride <= L.istride[0]);
}

// MappingOps.checkDimensions (MappingOps.java:122-138, MappingOps.cpp:73-91)
static void check_dimensions(const int32_t *dims, const int32_t *strides, int nd, int64_t len) {
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

static thread_local Scratch t_tmp[4];

// ---- warp / block primitives --------------------------------------------------------------------------

template <int OP>
__device__ __forceinline__ double rr_identity() {
    return OP == RR_PROD ? 1.0 : OP == RR_MAX ? -DBL_MAX : OP == RR_MIN ? DBL_MAX : 0.0;
}

template <int OP>
__device__ __forceinline__ double rr_combine(double a, double b) {
    return OP == RR_PROD ? a * b : OP == RR_MAX ? java_max(a, b) : OP == RR_MIN ? java_min(a, b) : a + b;
}

// ---- rrOp -----------------------------------------------------------------------------------------------
// One reduce step folds ONE dimension. RR_VAR is the reference's two-pass form (mean, then mean of squared
// differences: DimensionOpsReduce.cpp:231-258) run as two folds; SQ != 0 selects the second.

// Fold of points start, start + step, ... of one line with 8 independent accumulators: 8 loads in flight per
// thread instead of one dependent chain (the folds are latency-bound otherwise).
template <int OP, int SQ>
__device__ __forceinline__ double fold_line(const double *__restrict__ in, int64_t ioff, int64_t lstride, int64_t len,
        int start, int step, double mu) {
    // sums / products: 8 accumulators; MAX / MIN: 8 loads in flight but 2 accumulators (their fold is ~8 instructions
    // per point and 8 live accumulators cost the occupancy that hides the loads)
    constexpr int A = (OP == RR_MAX || OP == RR_MIN) ? 2 : 8;
    double acc[A];
    bool nan = false;
#pragma unroll
    for (int e = 0; e < A; e++) {
        acc[e] = rr_identity<OP>();
    }
    int64_t j = start;
    for (; j + 7 * (int64_t) step < len; j += 8 * (int64_t) step) {
        double x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            x[e] = in[ioff + (j + (int64_t) e * step) * lstride];
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
            if (SQ) {
                x[e] = (x[e] - mu) * (x[e] - mu);
            }
            if (OP == RR_MAX || OP == RR_MIN) {
                // Math.max / Math.min return NaN if either operand is NaN: fold with the NaN-ignoring fmax / fmin and
                // remember whether a NaN went by (one compare per point instead of java_max's three selects)
                nan |= x[e] != x[e];
                acc[e % A] = OP == RR_MAX ? fmax(acc[e % A], x[e]) : fmin(acc[e % A], x[e]);
            } else {
                acc[e % A] = rr_combine<OP>(acc[e % A], x[e]);
            }
        }
    }
    for (; j < len; j += step) {
        double x = in[ioff + j * lstride];
        if (SQ) {
            x = (x - mu) * (x - mu);
        }
        acc[0] = rr_combine<OP>(acc[0], x);
    }
#pragma unroll
    for (int w = A / 2; w > 0; w >>= 1) {
#pragma unroll
        for (int e = 0; e < w; e++) {
            acc[e] = rr_combine<OP>(acc[e], acc[e + w]);
        }
    }
    if ((OP == RR_MAX || OP == RR_MIN) && nan) {
        acc[0] = __longlong_as_double(0x7ff8000000000000LL);
    }
    return acc[0];
}

template <int OP, int SQ>
__global__ void __launch_bounds__(kThreads) reduce_lines_warp_kernel(const double *__restrict__ in, double *out,
        const double *mean, Lines L, double divisor) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t m = (((int64_t) blockIdx.x * blockDim.x) + threadIdx.x) >> 5; m < L.n_lines; m += warps) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        const double mu = SQ ? mean[ooff] : 0.0;
        double acc = fold_line<OP, SQ>(in, ioff, L.ilstride, L.len, lane, 32, mu);
        for (int off = 16; off > 0; off >>= 1) {
            acc = rr_combine<OP>(acc, __shfl_down_sync(0xffffffffu, acc, off));
        }
        if (lane == 0) {
            out[ooff] = divisor != 0.0 ? acc / divisor : acc;
        }
    }
}

template <int OP, int SQ>
__global__ void __launch_bounds__(kThreads) reduce_lines_block_kernel(const double *__restrict__ in, double *out,
        const double *mean, Lines L, double divisor) {
    __shared__ double part[kThreads / 32];
    for (int64_t m = blockIdx.x; m < L.n_lines; m += gridDim.x) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        const double mu = SQ ? mean[ooff] : 0.0;
        double acc = fold_line<OP, SQ>(in, ioff, L.ilstride, L.len, threadIdx.x, kThreads, mu);
        for (int off = 16; off > 0; off >>= 1) {
            acc = rr_combine<OP>(acc, __shfl_down_sync(0xffffffffu, acc, off));
        }
        if ((threadIdx.x & 31) == 0) {
            part[threadIdx.x >> 5] = acc;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            acc = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : rr_identity<OP>();
            for (int off = 16; off > 0; off >>= 1) {
                acc = rr_combine<OP>(acc, __shfl_down_sync(0xffffffffu, acc, off));
            }
            if (threadIdx.x == 0) {
                out[ooff] = divisor != 0.0 ? acc / divisor : acc;
            }
        }
        __syncthreads();
    }
}

// thread per (line, split): adjacent threads on adjacent lines. splits > 1 writes partials[split][line].
template <int OP, int SQ>
__global__ void __launch_bounds__(kThreads) reduce_lines_thread_kernel(const double *__restrict__ in, double *out,
        const double *mean, Lines L, double divisor, int splits, double *partials) {
    const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int sp = blockIdx.y;
    if (m >= L.n_lines) {
        return;
    }
    int64_t ioff, ooff;
    line_base(L, m, ioff, ooff);
    const int64_t chunk = (L.len + splits - 1) / splits;
    const int64_t lo = sp * chunk, hi = lo + chunk < L.len ? lo + chunk : L.len;
    const double mu = SQ ? mean[ooff] : 0.0;
    double acc = rr_identity<OP>();
    bool nan = false;
    auto take = [&](double x) {
        if (SQ) {
            x = (x - mu) * (x - mu);
        }
        if (OP == RR_MAX || OP == RR_MIN) {  // NaN-ignoring fold + NaN flag, as in fold_line
            nan |= x != x;
            acc = OP == RR_MAX ? fmax(acc, x) : fmin(acc, x);
        } else {
            acc = rr_combine<OP>(acc, x);
        }
    };
    int64_t j = lo;
    for (; j + 7 < hi; j += 8) {  // eight loads in flight, folded in the reference's order
        double x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            x[e] = in[ioff + (j + e) * L.ilstride];
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
            take(x[e]);
        }
    }
    for (; j < hi; j++) {
        take(in[ioff + j * L.ilstride]);
    }
    if ((OP == RR_MAX || OP == RR_MIN) && nan) {
        acc = __longlong_as_double(0x7ff8000000000000LL);
    }
    if (splits == 1) {
        out[ooff] = divisor != 0.0 ? acc / divisor : acc;
    } else {
        partials[(int64_t) sp * L.n_lines + m] = acc;
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) reduce_combine_kernel(const double *partials, double *out, Lines L,
        double divisor, int splits) {
    const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= L.n_lines) {
        return;
    }
    int64_t ioff, ooff;
    line_base(L, m, ioff, ooff);
    double acc = rr_identity<OP>();
    for (int sp = 0; sp < splits; sp++) {
        acc = rr_combine<OP>(acc, partials[(int64_t) sp * L.n_lines + m]);
    }
    out[ooff] = divisor != 0.0 ? acc / divisor : acc;
}

template <int OP, int SQ>
static void reduce_step(const double *in, double *out, const double *mean, const Lines &L, double divisor,
        cudaStream_t s) {
    if (L.n_lines == 0) {
        return;
    }
    if (line_is_fastest(L)) {
        // long lines: a CTA each, except plain sums / products over many lines, where a warp per line (no barrier)
        // measured faster at 8192^2 (0.095 vs 0.109 ms); MAX / MIN / squared-difference folds measured the other way
        const bool warp_each = L.n_lines >= (int64_t) sm_count() * 32 && (OP == RR_SUM || OP == RR_PROD) && !SQ;
        if (L.len >= 2048 && !warp_each) {
            int grid = (int) std::min<int64_t>(L.n_lines, (int64_t) sm_count() * 16);
            reduce_lines_block_kernel<OP, SQ><<<grid, kThreads, 0, s>>>(in, out, mean, L, divisor);
        } else {
            int64_t blocks = (L.n_lines * 32 + kThreads - 1) / kThreads;
            int grid = (int) std::min<int64_t>(blocks, (int64_t) sm_count() * 16);
            reduce_lines_warp_kernel<OP, SQ><<<grid, kThreads, 0, s>>>(in, out, mean, L, divisor);
        }
        check_launch("reduce_lines_kernel");
        return;
    }
    // thread per line; split long lines so that the grid still fills the machine
    const int64_t target = (int64_t) sm_count() * 2048;
    int splits = 1;
    if (L.n_lines < target && L.len >= 128) {
        splits = (int) std::min<int64_t>((target + L.n_lines - 1) / L.n_lines, L.len / 64);
        splits = std::max(1, std::min(splits, 256));
    }
    dim3 grid((unsigned) ((L.n_lines + kThreads - 1) / kThreads), (unsigned) splits);
    double *partials = splits > 1 ? static_cast<double *>(t_tmp[3].get((int64_t) splits * L.n_lines * 8)) : nullptr;
    reduce_lines_thread_kernel<OP, SQ><<<grid, kThreads, 0, s>>>(in, out, mean, L, divisor, splits, partials);
    check_launch("reduce_lines_thread_kernel");
    if (splits > 1) {
        reduce_combine_kernel<OP><<<grid.x, kThreads, 0, s>>>(partials, out, L, divisor, splits);
        check_launch("reduce_combine_kernel");
    }
}

static void reduce_dim(int type, const double *in, double *out, double *mean_buf, const Lines &L, cudaStream_t s) {
    switch (type) {
    case RR_SUM: reduce_step<RR_SUM, 0>(in, out, nullptr, L, 0.0, s); break;
    case RR_PROD: reduce_step<RR_PROD, 0>(in, out, nullptr, L, 0.0, s); break;
    case RR_MAX: reduce_step<RR_MAX, 0>(in, out, nullptr, L, 0.0, s); break;
    case RR_MIN: reduce_step<RR_MIN, 0>(in, out, nullptr, L, 0.0, s); break;
    default: {
        // mean into `out` itself (same addressing), then the squared differences against it
        reduce_step<RR_SUM, 0>(in, mean_buf, nullptr, L, (double) L.len, s);
        reduce_step<RR_SUM, 1>(in, out, mean_buf, L, (double) L.len, s);
    }
    }
}

// dst[o(idx)] = src[i(idx)] over every index: the reference's final double-indexed gather
// (DimensionOpsReduce.cpp:166-171) and rdOp's initial memcpy when layouts differ.
template <class T>
__global__ void __launch_bounds__(kThreads) strided_copy_kernel(const T *__restrict__ in, T *out, Lines L) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; m < L.n_lines; m += stride) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        out[ooff] = in[ioff];
    }
}

// ---- rdOp -----------------------------------------------------------------------------------------------

template <int OP>
__device__ __forceinline__ double rd_combine(double a, double b) {
    return OP == RD_SUM ? a + b : a * b;
}

// CTA per line, lanes along the line; rounds of blockDim points with a running carry.
// Each thread scans kScanItems consecutive points in registers, the thread totals are scanned with shuffles
// and one shared-memory hop per round of 256 * kScanItems points, and a running carry links the rounds.
constexpr int kScanItems = 8;

template <int OP>
__global__ void __launch_bounds__(kThreads) scan_lines_block_kernel(const double *in, double *out, Lines L) {
    __shared__ double warp_tot[kThreads / 32];
    __shared__ double carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double ident = OP == RD_SUM ? 0.0 : 1.0;
    for (int64_t m = blockIdx.x; m < L.n_lines; m += gridDim.x) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        double carry = ident;
        // A round covers 256 * kScanItems points; warp w owns the contiguous run of 32 * kScanItems points
        // starting at base + w*32*kScanItems and walks it in kScanItems coalesced 32-point steps, scanning each
        // step with shuffles and chaining the steps through lane 31 -- no shared memory inside a warp's run.
        for (int64_t base = 0; base < L.len; base += (int64_t) blockDim.x * kScanItems) {
            const int64_t w0 = base + (int64_t) warp * 32 * kScanItems;
            double x[kScanItems];
            double run = ident;  // inclusive total of this warp's run so far
#pragma unroll
            for (int e = 0; e < kScanItems; e++) {
                const int64_t j = w0 + e * 32 + lane;
                double v = j < L.len ? in[ioff + j * L.ilstride] : ident;
                for (int off = 1; off < 32; off <<= 1) {
                    double y = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) {
                        v = rd_combine<OP>(y, v);
                    }
                }
                v = rd_combine<OP>(run, v);
                run = __shfl_sync(0xffffffffu, v, 31);
                x[e] = v;
            }
            if (lane == 0) {
                warp_tot[warp] = run;
            }
            __syncthreads();
            double pre = carry;
            for (int w = 0; w < warp; w++) {
                pre = rd_combine<OP>(pre, warp_tot[w]);
            }
#pragma unroll
            for (int e = 0; e < kScanItems; e++) {
                const int64_t j = w0 + e * 32 + lane;
                if (j < L.len) {
                    out[ooff + j * L.olstride] = rd_combine<OP>(pre, x[e]);
                }
            }
            if (threadIdx.x == blockDim.x - 1) {
                carry_s = rd_combine<OP>(pre, run);
            }
            __syncthreads();
            carry = carry_s;
        }
    }
}

// Warp per line: short lines, or long ones when there are enough lines to fill the machine with warps (no CTA
// barrier at all). kWarpSteps coalesced 32-point steps are loaded up front (2 KB in flight per warp), then scanned
// one after another with shuffles, chained through lane 31.
constexpr int kWarpSteps = 8;

template <int OP>
__global__ void __launch_bounds__(kThreads) scan_lines_warp_kernel(const double *in, double *out, Lines L) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    const double ident = OP == RD_SUM ? 0.0 : 1.0;
    for (int64_t m = (((int64_t) blockIdx.x * blockDim.x) + threadIdx.x) >> 5; m < L.n_lines; m += warps) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        double carry = ident;
        for (int64_t base = 0; base < L.len; base += 32 * kWarpSteps) {
            double x[kWarpSteps];
#pragma unroll
            for (int e = 0; e < kWarpSteps; e++) {
                const int64_t j = base + e * 32 + lane;
                x[e] = j < L.len ? in[ioff + j * L.ilstride] : ident;
            }
#pragma unroll
            for (int e = 0; e < kWarpSteps; e++) {
                const int64_t j = base + e * 32 + lane;
                double v = x[e];
                for (int off = 1; off < 32; off <<= 1) {
                    const double y = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) {
                        v = rd_combine<OP>(y, v);
                    }
                }
                v = rd_combine<OP>(carry, v);
                if (j < L.len) {
                    out[ooff + j * L.olstride] = v;
                }
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
    }
}

// Thread per line, sequential along it (the reference's order exactly). Long lines are cut into `splits`
// chunks: pass 1 (PHASE 0) folds each chunk, pass 2 (PHASE 1) rescans each chunk from the fold of the chunks
// before it -- 1.5x the traffic of a single pass, in exchange for splits-times the parallelism.
template <int OP, int PHASE>
__global__ void __launch_bounds__(kThreads) scan_lines_thread_kernel(const double *in, double *out, Lines L,
        int splits, double *partials) {
    const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int sp = blockIdx.y;
    if (m >= L.n_lines) {
        return;
    }
    int64_t ioff, ooff;
    line_base(L, m, ioff, ooff);
    const int64_t chunk = (L.len + splits - 1) / splits;
    const int64_t lo = sp * chunk, hi = lo + chunk < L.len ? lo + chunk : L.len;
    double acc = OP == RD_SUM ? 0.0 : 1.0;
    if (PHASE == 0) {
        for (int64_t j = lo; j < hi; j++) {
            acc = rd_combine<OP>(acc, in[ioff + j * L.ilstride]);
        }
        partials[(int64_t) sp * L.n_lines + m] = acc;
    } else {
        for (int q = 0; q < sp && splits > 1; q++) {
            acc = rd_combine<OP>(acc, partials[(int64_t) q * L.n_lines + m]);
        }
        for (int64_t j = lo; j < hi; j++) {
            acc = rd_combine<OP>(acc, in[ioff + j * L.ilstride]);
            out[ooff + j * L.olstride] = acc;
        }
    }
}

// Lines that are NOT along the fastest dimension, many of them (an image's columns): a CTA owns a strip of 32
// adjacent lines (lane = line, so every row of the strip is one coalesced 256-byte access) and walks down it in
// blocks of 8 warps x kStripRows rows (32: measured 0.92 of HBM at 8192^2 against 0.80 for 16): each warp scans its kStripRows rows in registers, the per-warp column
// totals meet in shared memory, and a per-lane carry links the blocks. One pass: 8 B read + 8 B written per point.
constexpr int kStripRows = 32;

template <int OP>
__global__ void __launch_bounds__(kThreads) scan_strip_kernel(const double *in, double *out, Lines L) {
    constexpr int NW = kThreads / 32;
    __shared__ double tot[2][NW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double ident = OP == RD_SUM ? 0.0 : 1.0;
    const int64_t strips = (L.n_lines + 31) / 32;
    for (int64_t sidx = blockIdx.x; sidx < strips; sidx += gridDim.x) {
        const int64_t m = sidx * 32 + lane;
        const bool live = m < L.n_lines;
        int64_t ioff = 0, ooff = 0;
        if (live) {
            line_base(L, m, ioff, ooff);
        }
        double carry = ident;
        int buf = 0;
        for (int64_t base = 0; base < L.len; base += (int64_t) NW * kStripRows, buf ^= 1) {
            const int64_t r0 = base + (int64_t) warp * kStripRows;
            double x[kStripRows];
#pragma unroll
            for (int e = 0; e < kStripRows; e++) {
                x[e] = (live && r0 + e < L.len) ? in[ioff + (r0 + e) * L.ilstride] : ident;
            }
#pragma unroll
            for (int e = 1; e < kStripRows; e++) {
                x[e] = rd_combine<OP>(x[e - 1], x[e]);
            }
            tot[buf][warp][lane] = x[kStripRows - 1];
            __syncthreads();  // tot[buf] is rewritten two blocks later, after the next barrier: no second barrier
            double pre = carry;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const double t = tot[buf][w][lane];
                pre = w < warp ? rd_combine<OP>(pre, t) : pre;
                carry = rd_combine<OP>(carry, t);
            }
#pragma unroll
            for (int e = 0; e < kStripRows; e++) {
                if (live && r0 + e < L.len) {
                    out[ooff + (r0 + e) * L.olstride] = rd_combine<OP>(pre, x[e]);
                }
            }
        }
        __syncthreads();
    }
}

template <int OP>
static void scan_dim(const double *in, double *out, const Lines &L, cudaStream_t s) {
    if (L.n_lines == 0 || L.len == 0) {
        return;
    }
    if (line_is_fastest(L)) {
        if (L.len > 256 && L.n_lines < (int64_t) sm_count() * 32) {
            int grid = (int) std::min<int64_t>(L.n_lines, (int64_t) sm_count() * 16);
            scan_lines_block_kernel<OP><<<grid, kThreads, 0, s>>>(in, out, L);
        } else {
            int64_t blocks = (L.n_lines * 32 + kThreads - 1) / kThreads;
            int grid = (int) std::min<int64_t>(blocks, (int64_t) sm_count() * 16);
            scan_lines_warp_kernel<OP><<<grid, kThreads, 0, s>>>(in, out, L);
        }
        check_launch("scan_lines_kernel");
        return;
    }
    if (L.n_lines >= (int64_t) sm_count() * 32 && L.len >= 64) {
        const int grid = (int) std::min<int64_t>((L.n_lines + 31) / 32, (int64_t) sm_count() * 8);
        scan_strip_kernel<OP><<<grid, kThreads, 0, s>>>(in, out, L);
        check_launch("scan_strip_kernel");
        return;
    }
    const int64_t target = (int64_t) sm_count() * 2048;
    int splits = 1;
    if (L.n_lines < target && L.len >= 256) {
        splits = (int) std::min<int64_t>((target + L.n_lines - 1) / L.n_lines, L.len / 64);
        splits = std::max(1, std::min(splits, 64));
    }
    dim3 grid((unsigned) ((L.n_lines + kThreads - 1) / kThreads), (unsigned) splits);
    double *partials = nullptr;
    if (splits > 1) {
        partials = static_cast<double *>(t_tmp[3].get((int64_t) splits * L.n_lines * 8));
        scan_lines_thread_kernel<OP, 0><<<grid, kThreads, 0, s>>>(in, out, L, splits, partials);
        check_launch("scan_lines_thread_kernel");
    }
    scan_lines_thread_kernel<OP, 1><<<grid, kThreads, 0, s>>>(in, out, L, splits, partials);
    check_launch("scan_lines_thread_kernel");
}

// ---- riOp: predicates and ordered compaction --------------------------------------------------------------

template <int OP>
__device__ __forceinline__ bool ri_pred(double x, double ext) {
    return OP == RI_ZERO ? x == 0.0 : OP == RI_GZERO ? x > 0.0 : OP == RI_LZERO ? x < 0.0 : x == ext;
}

// CTA per line: (1) one coalesced read of the line (cached in shared memory when it fits) and, for MAX / MIN,
// the extreme; (2) every warp counts the matches in its contiguous run of the line; (3) after one prefix over
// the warp counts, every warp rewrites its matches' positions in ascending order behind the earlier warps',
// then the tail is filled with -1 (DimensionOps.java:186-199). Three CTA barriers per line in total.
template <int OP, bool CACHE>
__global__ void __launch_bounds__(kThreads) index_lines_block_kernel(const double *__restrict__ in, int32_t *out,
        Lines L) {
    extern __shared__ double sline[];
    __shared__ double part[kThreads / 32];
    __shared__ int warp_cnt[kThreads / 32];
    constexpr int NW = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunk = ((L.len + NW - 1) / NW + 31) / 32 * 32;  // per-warp run, a multiple of 32
    for (int64_t m = blockIdx.x; m < L.n_lines; m += gridDim.x) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        double ext = 0.0;
        if (CACHE || OP == RI_MAX || OP == RI_MIN) {
            double acc = OP == RI_MAX ? -DBL_MAX : DBL_MAX;
            for (int64_t j = threadIdx.x; j < L.len; j += blockDim.x) {
                const double x = in[ioff + j * L.ilstride];
                if (CACHE) {
                    sline[j] = x;
                }
                if (OP == RI_MAX) {
                    acc = java_max(acc, x);
                } else if (OP == RI_MIN) {
                    acc = java_min(acc, x);
                }
            }
            if (OP == RI_MAX || OP == RI_MIN) {
                for (int off = 16; off > 0; off >>= 1) {
                    const double o = __shfl_down_sync(0xffffffffu, acc, off);
                    acc = OP == RI_MAX ? java_max(acc, o) : java_min(acc, o);
                }
                if (lane == 0) {
                    part[warp] = acc;
                }
            }
            __syncthreads();
            if (OP == RI_MAX || OP == RI_MIN) {
                ext = part[0];
                for (int w = 1; w < NW; w++) {
                    ext = OP == RI_MAX ? java_max(ext, part[w]) : java_min(ext, part[w]);
                }
            }
        }
        const int64_t lo = warp * chunk, hi = lo + chunk < L.len ? lo + chunk : L.len;
        int cnt = 0;
        for (int64_t j = lo + lane; j < lo + chunk; j += 32) {
            const bool hit = j < hi && ri_pred<OP>(CACHE ? sline[j] : in[ioff + j * L.ilstride], ext);
            cnt += __popc(__ballot_sync(0xffffffffu, hit));
        }
        if (lane == 0) {
            warp_cnt[warp] = cnt;
        }
        __syncthreads();
        int pos = 0, total = 0;
        for (int w = 0; w < NW; w++) {
            pos += w < warp ? warp_cnt[w] : 0;
            total += warp_cnt[w];
        }
        for (int64_t j = lo + lane; j < lo + chunk; j += 32) {
            const bool hit = j < hi && ri_pred<OP>(CACHE ? sline[j] : in[ioff + j * L.ilstride], ext);
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                out[ooff + (int64_t) (pos + __popc(ballot & ((1u << lane) - 1u))) * L.olstride] = (int32_t) j;
            }
            pos += __popc(ballot);
        }
        for (int64_t j = total + threadIdx.x; j < L.len; j += blockDim.x) {
            out[ooff + j * L.olstride] = -1;
        }
        __syncthreads();
    }
}

// thread per line, the reference's loops verbatim in structure (DimensionOpsIndex.cpp:128-254)
template <int OP>
__global__ void __launch_bounds__(kThreads) index_lines_thread_kernel(const double *__restrict__ in, int32_t *out,
        Lines L) {
    const int64_t m = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= L.n_lines) {
        return;
    }
    int64_t ioff, ooff;
    line_base(L, m, ioff, ooff);
    double ext = 0.0;
    if (OP == RI_MAX || OP == RI_MIN) {
        ext = OP == RI_MAX ? -DBL_MAX : DBL_MAX;
        for (int64_t j = 0; j < L.len; j++) {
            double x = in[ioff + j * L.ilstride];
            ext = OP == RI_MAX ? java_max(ext, x) : java_min(ext, x);
        }
    }
    int64_t count = 0;
    for (int64_t j = 0; j < L.len; j++) {
        if (ri_pred<OP>(in[ioff + j * L.ilstride], ext)) {
            out[ooff + count * L.olstride] = (int32_t) j;
            count++;
        }
    }
    for (int64_t j = count; j < L.len; j++) {
        out[ooff + j * L.olstride] = -1;
    }
}

// Warp per line, for many lines along the fastest dimension: no CTA barrier, no shared memory. Pass 1 (MAX / MIN
// only) folds the line with 8 loads in flight per lane; pass 2 re-reads it (it is 16*len bytes per warp and comes
// back from L2), ballots the matches of every 32-point step and writes their positions in ascending order behind
// a running count; the tail is filled with -1 (DimensionOps.java:186-199).
template <int OP>
__global__ void __launch_bounds__(kThreads) index_lines_warp_kernel(const double *__restrict__ in, int32_t *out,
        Lines L) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    for (int64_t m = (((int64_t) blockIdx.x * blockDim.x) + threadIdx.x) >> 5; m < L.n_lines; m += warps) {
        int64_t ioff, ooff;
        line_base(L, m, ioff, ooff);
        double ext = 0.0;
        int pos = 0;
        bool need_scan = true;
        if (OP == RI_MAX || OP == RI_MIN) {
            // Every lane tracks, per accumulator, its extreme, where it first saw it and how often. If the line's
            // extreme occurs once (the usual case for real-valued data) its position is known after this one pass.
            double best[8];
            int first[8], cnt[8];
            bool nan = false;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                best[e] = OP == RI_MAX ? -DBL_MAX : DBL_MAX;
                first[e] = 0x7fffffff;
                cnt[e] = 0;
            }
            auto better = [](double x, double y) { return OP == RI_MAX ? x > y : x < y; };
            auto take = [&](int e, double x, int j) {
                nan |= x != x;
                if (better(x, best[e])) {
                    best[e] = x;
                    first[e] = j;
                    cnt[e] = 1;
                } else if (x == best[e]) {
                    first[e] = first[e] < j ? first[e] : j;
                    cnt[e]++;
                }
            };
            int64_t j = lane;
            for (; j + 7 * 32 < L.len; j += 8 * 32) {
                double x[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    x[e] = in[ioff + (j + e * 32) * L.ilstride];
                }
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    take(e, x[e], (int) j + e * 32);
                }
            }
            for (; j < L.len; j += 32) {
                take(0, in[ioff + j * L.ilstride], (int) j);
            }
            auto merge = [&](double &b0, int &f0, int &c0, double b1, int f1, int c1) {
                if (better(b1, b0)) {
                    b0 = b1;
                    f0 = f1;
                    c0 = c1;
                } else if (b1 == b0) {
                    f0 = f0 < f1 ? f0 : f1;
                    c0 += c1;
                }
            };
#pragma unroll
            for (int e = 1; e < 8; e++) {
                merge(best[0], first[0], cnt[0], best[e], first[e], cnt[e]);
            }
            for (int off = 16; off > 0; off >>= 1) {
                const double b1 = __shfl_xor_sync(0xffffffffu, best[0], off);
                const int f1 = __shfl_xor_sync(0xffffffffu, first[0], off);
                const int c1 = __shfl_xor_sync(0xffffffffu, cnt[0], off);
                merge(best[0], first[0], cnt[0], b1, f1, c1);
            }
            ext = best[0];
            if (__any_sync(0xffffffffu, nan)) {
                need_scan = false;  // Math.max / Math.min give NaN, and nothing equals NaN: no positions
            } else if (cnt[0] == 1) {
                if (lane == 0) {
                    out[ooff] = first[0];
                }
                pos = 1;
                need_scan = false;
            } else if (cnt[0] == 0) {
                // nothing beat the start value -/+Double.MAX_VALUE (the line holds only -/+infinity): positions are
                // those equal to the start value, as the reference's fold leaves it (Arithmetic.java:54,72) -> none
                need_scan = true;
            }
        }
        for (int64_t base = 0; need_scan && base < L.len; base += 8 * 32) {
            double x[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int64_t j = base + e * 32 + lane;
                x[e] = j < L.len ? in[ioff + j * L.ilstride] : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int64_t j = base + e * 32 + lane;
                const bool hit = j < L.len && ri_pred<OP>(x[e], ext);
                const unsigned ballot = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    out[ooff + (int64_t) (pos + __popc(ballot & ((1u << lane) - 1u))) * L.olstride] = (int32_t) j;
                }
                pos += __popc(ballot);
            }
        }
        for (int64_t j = pos + lane; j < L.len; j += 32) {
            out[ooff + j * L.olstride] = -1;
        }
    }
}

template <int OP>
static void index_dim(const double *in, int32_t *out, const Lines &L, cudaStream_t s) {
    if (L.n_lines == 0 || L.len == 0) {
        return;
    }
    if (line_is_fastest(L) && L.len >= 64 && L.n_lines >= (int64_t) sm_count() * 32 && L.len <= 2147483647LL) {
        const int64_t blocks = (L.n_lines * 32 + kThreads - 1) / kThreads;
        index_lines_warp_kernel<OP><<<(int) std::min<int64_t>(blocks, (int64_t) sm_count() * 16), kThreads, 0, s>>>(in, out, L);
    } else if (line_is_fastest(L) && L.len >= 64) {
        int grid = (int) std::min<int64_t>(L.n_lines, (int64_t) sm_count() * 16);
        constexpr int kCacheMax = 12288;  // 96 KB of shared memory: two lines resident per SM
        if (L.len <= kCacheMax && (OP == RI_MAX || OP == RI_MIN)) {
            static std::once_flag once[16];
            int dev = 0;
            SSTC_CUDA(cudaGetDevice(&dev));
            std::call_once(once[dev & 15], [] {
                SSTC_CUDA(cudaFuncSetAttribute(index_lines_block_kernel<OP, true>,
                        cudaFuncAttributeMaxDynamicSharedMemorySize, kCacheMax * 8));
            });
            index_lines_block_kernel<OP, true><<<grid, kThreads, (size_t) L.len * 8, s>>>(in, out, L);
        } else {
            index_lines_block_kernel<OP, false><<<grid, kThreads, 0, s>>>(in, out, L);
        }
    } else {
        index_lines_thread_kernel<OP><<<(unsigned) ((L.n_lines + kThreads - 1) / kThreads), kThreads, 0, s>>>(in, out, L);
    }
    check_launch("index_lines_kernel");
}

// dim == -1: 0/1 mask over the physical array (DimensionOpsIndex.cpp:141-146 etc.)
template <int OP>
__global__ void __launch_bounds__(kThreads) index_mask_kernel(const double *__restrict__ in, int32_t *out, int64_t n,
        const double *ext) {
    const double e = (OP == RI_MAX || OP == RI_MIN) ? ext[0] : 0.0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = ri_pred<OP>(in[i], e) ? 1 : 0;
    }
}

// ---- segmented bitonic sort of (key, payload) pairs ---------------------------------------------------------
// Keys are order-preserving 64-bit images of doubles; ties break on the payload (the original position), so
// the result is the STABLE ascending order -- what the Java provider's Arrays.sort of PermutationEntry gives
// (DimensionOps.java:362-377, PermutationEntry.java:94-140). The reference's C++ uses std::sort, which is not
// stable; the two agree whenever a line has no repeated values.

__device__ __forceinline__ uint64_t double_key(double x) {
    uint64_t b = (uint64_t) __double_as_longlong(x);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

__device__ __forceinline__ double key_double(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long) b);
}

// work layout: [line][npad]; pads carry the largest key so they sink to the end of their segment
__global__ void __launch_bounds__(kThreads) sort_gather_kernel(const double *__restrict__ in, uint64_t *keys,
        int32_t *pay, Lines L, int64_t npad) {
    const int64_t total = L.n_lines * npad;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t m = i / npad, j = i - m * npad;
        if (j < L.len) {
            int64_t ioff, ooff;
            line_base(L, m, ioff, ooff);
            keys[i] = double_key(in[ioff + j * L.ilstride]);
            pay[i] = (int32_t) j;
        } else {
            keys[i] = ~0ULL;
            pay[i] = INT_MAX;
        }
    }
}

__global__ void __launch_bounds__(kThreads) sort_scatter_kernel(double *in, int32_t *out, const uint64_t *keys,
        const int32_t *pay, Lines L, int64_t npad) {
    const int64_t total = L.n_lines * npad;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t m = i / npad, j = i - m * npad;
        if (j < L.len) {
            int64_t ioff, ooff;
            line_base(L, m, ioff, ooff);
            in[ioff + j * L.ilstride] = key_double(keys[i]);  // RI_SORT sorts the source in place
            out[ooff + j * L.olstride] = pay[i];
        }
    }
}

__device__ __forceinline__ bool pair_less(uint64_t ka, int32_t pa, uint64_t kb, int32_t pb) {
    return ka < kb || (ka == kb && pa < pb);
}

// all (k, j) steps with j < TILE/2... are done in shared memory on TILE consecutive pairs
constexpr int kSortTile = 2048;

__global__ void __launch_bounds__(kThreads) bitonic_global_kernel(uint64_t *keys, int32_t *pay, int64_t total,
        int64_t k, int64_t j, int64_t seg) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < (total >> 1); t += stride) {
        // t-th pair (i, i ^ j) with i having bit j clear
        const int64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int64_t p = i | j;
        const bool up = k == seg || (i & k) == 0;  // the last merge of every segment is ascending
        const uint64_t ka = keys[i], kb = keys[p];
        const int32_t pa = pay[i], pb = pay[p];
        if (pair_less(kb, pb, ka, pa) == up) {
            keys[i] = kb;
            keys[p] = ka;
            pay[i] = pb;
            pay[p] = pa;
        }
    }
}

// Runs, for one k, every step j = jstart, jstart/2, ..., 1 (jstart < kSortTile) inside shared memory; with
// FULL it runs the whole network up to k = kSortTile first (the opening phase of the sort).
template <bool FULL>
__global__ void __launch_bounds__(kThreads) bitonic_tile_kernel(uint64_t *keys, int32_t *pay, int64_t total,
        int64_t k_outer, int64_t jstart, int64_t seg) {
    __shared__ uint64_t sk[kSortTile];
    __shared__ int32_t sp[kSortTile];
    for (int64_t tile = blockIdx.x; tile * kSortTile < total; tile += gridDim.x) {
        const int64_t base = tile * kSortTile;
        for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) {
            sk[t] = keys[base + t];
            sp[t] = pay[base + t];
        }
        __syncthreads();
        for (int64_t k = FULL ? 2 : k_outer; k <= (FULL ? (int64_t) kSortTile : k_outer); k <<= 1) {
            for (int64_t j = FULL ? (k >> 1) : jstart; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < kSortTile / 2; t += blockDim.x) {
                    const int i = (int) (((t & ~(j - 1)) << 1) | (t & (j - 1)));
                    const int p = i | (int) j;
                    const bool up = k == seg || ((base + i) & k) == 0;
                    const uint64_t ka = sk[i], kb = sk[p];
                    const int32_t pa = sp[i], pb = sp[p];
                    if (pair_less(kb, pb, ka, pa) == up) {
                        sk[i] = kb;
                        sk[p] = ka;
                        sp[i] = pb;
                        sp[p] = pa;
                    }
                }
                __syncthreads();
            }
        }
        for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) {
            keys[base + t] = sk[t];
            pay[base + t] = sp[t];
        }
        __syncthreads();
    }
}

static int64_t next_pow2(int64_t n) {
    int64_t p = 1;
    while (p < n) {
        p <<= 1;
    }
    return p;
}

// Sorts every aligned segment of `seg` (a power of two) pairs ascending; `total` is a multiple of seg.
static void bitonic_sort_segments(uint64_t *keys, int32_t *pay, int64_t total, int64_t seg, cudaStream_t s) {
    if (seg < 2 || total == 0) {
        return;
    }
    const int cap = sm_count() * 8;
    if (total % kSortTile == 0 && seg >= 2) {
        const int64_t tiles = total / kSortTile;
        const int grid = (int) std::min<int64_t>(tiles, cap);
        // opening: the full network up to k = min(seg, tile); a segment-local network never crosses a tile
        // boundary when seg <= tile, and tiles are aligned, so this is exact for both cases
        if (seg <= kSortTile) {
            // run k = 2..seg with all j in shared memory: emulate with per-k calls
            for (int64_t k = 2; k <= seg; k <<= 1) {
                bitonic_tile_kernel<false><<<grid, kThreads, 0, s>>>(keys, pay, total, k, k >> 1, seg);
                check_launch("bitonic_tile_kernel");
            }
            return;
        }
        bitonic_tile_kernel<true><<<grid, kThreads, 0, s>>>(keys, pay, total, 0, 0, seg);
        check_launch("bitonic_tile_kernel");
        const int ggrid = (int) std::min<int64_t>((total / 2 + kThreads - 1) / kThreads, (int64_t) cap * 4);
        for (int64_t k = (int64_t) kSortTile << 1; k <= seg; k <<= 1) {
            int64_t j = k >> 1;
            for (; j >= kSortTile; j >>= 1) {
                bitonic_global_kernel<<<ggrid, kThreads, 0, s>>>(keys, pay, total, k, j, seg);
                check_launch("bitonic_global_kernel");
            }
            bitonic_tile_kernel<false><<<grid, kThreads, 0, s>>>(keys, pay, total, k, j, seg);
            check_launch("bitonic_tile_kernel");
        }
        return;
    }
    // small or unaligned totals: plain global steps
    const int ggrid = (int) std::max<int64_t>(1, std::min<int64_t>((total / 2 + kThreads - 1) / kThreads, (int64_t) cap * 4));
    for (int64_t k = 2; k <= seg; k <<= 1) {
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            bitonic_global_kernel<<<ggrid, kThreads, 0, s>>>(keys, pay, total, k, j, seg);
            check_launch("bitonic_global_kernel");
        }
    }
}

// the same network for other callers (array_sparse.cu): `total` pairs in aligned power-of-two segments of `seg`
void sort_pairs_segments(uint64_t *keys, int32_t *pay, int64_t total, int64_t seg, cudaStream_t s) {
    bitonic_sort_segments(keys, pay, total, seg, s);
}

static void sort_dim(double *in, int32_t *out, const Lines &L, cudaStream_t s) {
    if (L.n_lines == 0 || L.len == 0) {
        return;
    }
    const int64_t npad = next_pow2(L.len);
    const int64_t total = L.n_lines * npad;
    uint64_t *keys = static_cast<uint64_t *>(t_tmp[0].get(total * 8));
    int32_t *pay = static_cast<int32_t *>(t_tmp[1].get(total * 4));
    const int grid = stream_grid(total);
    sort_gather_kernel<<<grid, kThreads, 0, s>>>(in, keys, pay, L, npad);
    check_launch("sort_gather_kernel");
    bitonic_sort_segments(keys, pay, total, npad, s);
    sort_scatter_kernel<<<grid, kThreads, 0, s>>>(in, out, keys, pay, L, npad);
    check_launch("sort_scatter_kernel");
}

// ---- SHUFFLE: a uniformly random permutation = sort by random keys ------------------------------------------

__global__ void __launch_bounds__(kThreads) shuffle_keys_kernel(uint64_t *keys, int32_t *pay, int64_t n, int64_t npad,
        uint64_t key) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += stride) {
        keys[i] = i < n ? (splitmix64(key ^ splitmix64((uint64_t) i)) >> 1) : ~0ULL;
        pay[i] = i < n ? (int32_t) i : INT_MAX;
    }
}

template <class T>
__global__ void __launch_bounds__(kThreads) permute_kernel(const T *__restrict__ in, T *out, const int32_t *perm,
        int64_t n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = in[perm[i]];
    }
}

// ElementOps.java uShuffle / ElementOps.cpp:686-693 (Fisher-Yates on the host RNG): specified only as "a
// uniformly random permutation" (ArrayKernelTest.java:151-154).
void shuffle_dev(void *v, int64_t n, int elem_bytes, cudaStream_t s) {
    if (n < 2) {
        return;
    }
    if (n > INT_MAX) {
        throw std::runtime_error("Invalid arguments");
    }
    const int64_t npad = next_pow2(n);
    uint64_t *keys = static_cast<uint64_t *>(t_tmp[0].get(npad * 8));
    int32_t *pay = static_cast<int32_t *>(t_tmp[1].get(npad * 4));
    void *tmp = t_tmp[2].get(n * elem_bytes);
    const int grid = stream_grid(npad);
    shuffle_keys_kernel<<<grid, kThreads, 0, s>>>(keys, pay, n, npad, next_rng_key());
    check_launch("shuffle_keys_kernel");
    bitonic_sort_segments(keys, pay, npad, npad, s);
    if (elem_bytes == 8) {
        permute_kernel<double><<<grid, kThreads, 0, s>>>(static_cast<const double *>(v), static_cast<double *>(tmp), pay, n);
    } else if (elem_bytes == 16) {
        permute_kernel<double2><<<grid, kThreads, 0, s>>>(static_cast<const double2 *>(v), static_cast<double2 *>(tmp), pay,
                n);
    } else {
        permute_kernel<int32_t><<<grid, kThreads, 0, s>>>(static_cast<const int32_t *>(v), static_cast<int32_t *>(tmp), pay,
                n);
    }
    check_launch("permute_kernel");
    SSTC_CUDA(cudaMemcpyAsync(v, tmp, (size_t) (n * elem_bytes), cudaMemcpyDeviceToDevice, s));
}

static std::vector<int64_t> widen(const int32_t *p, int n) { return std::vector<int64_t>(p, p + n); }

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_rrop_dev(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, int32_t *opDims, int nop, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        if (type < RR_SUM || type > RR_VAR) {
            throw std::runtime_error("Operation type not recognized");
        }
        if (!sD || !sS || !dD || !dS || (nop > 0 && !opDims) || nd < 0 || nop < 0 || nsrc < 0 || ndst < 0 ||
                (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nd > kMaxDims) {
            throw std::runtime_error("Too many dimensions for the CUDA provider");
        }
        check_dimensions(sD, sS, nd, nsrc);
        check_dimensions(dD, dS, nd, ndst);
        std::sort(opDims, opDims + nop);  // the reference sorts the caller's array too (DimensionOpsReduce.cpp:98)
        for (int i = 1; i < nop; i++) {
            if (opDims[i - 1] == opDims[i]) {
                throw std::runtime_error("Duplicate operating dimensions are not allowed");
            }
        }
        int64_t acc = ndst;
        for (int i = 0; i < nop; i++) {
            const int dim = opDims[i];
            if (!(dim >= 0 && dim < nd)) {
                throw std::runtime_error("Invalid dimension");
            }
            if (dD[dim] > 1) {
                throw std::runtime_error("Operating dimensions must have singleton or zero length");
            }
            acc *= sD[dim];
        }
        if (acc != nsrc) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nsrc == 0) {
            return;
        }
        std::vector<int64_t> cur_d = widen(sD, nd), cur_s = widen(sS, nd), dst_s = widen(dS, nd);
        const double *cur = src;
        if (nop == 0) {
            Lines L = make_lines(cur_d, cur_s, dst_s, -1);
            strided_copy_kernel<double><<<stream_grid(L.n_lines), kThreads, 0, s>>>(cur, dst, L);
            check_launch("strided_copy_kernel");
            return;
        }
        int64_t remaining = nsrc;
        for (int i = 0; i < nop; i++) {
            const int dim = opDims[i];
            const bool last = i == nop - 1;
            std::vector<int64_t> out_d = cur_d;
            out_d[(size_t) dim] = 1;
            remaining /= cur_d[(size_t) dim];
            std::vector<int64_t> out_s = last ? dst_s : far_strides(out_d);
            double *out = last ? dst : static_cast<double *>(t_tmp[i & 1].get(remaining * 8));
            double *mean_buf = nullptr;
            if (type == RR_VAR) {
                // means are addressed like the outputs; give them a buffer with the output's extent
                int64_t extent = 1;
                for (int k = 0; k < nd; k++) {
                    extent += (out_d[(size_t) k] - 1) * out_s[(size_t) k];
                }
                mean_buf = static_cast<double *>(t_tmp[2].get(extent * 8));
            }
            Lines L = make_lines(cur_d, cur_s, out_s, dim);
            reduce_dim(type, cur, out, mean_buf, L, s);
            cur = out;
            cur_d = out_d;
            cur_s = out_s;
        }
    });
}

int sstc_rdop_dev(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, double *dst,
        int64_t ndst, const int32_t *opDims, int nop, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        if (type < RD_SUM || type > RD_PROD) {
            throw std::runtime_error("Operation type not recognized");
        }
        if (!sD || !sS || (nop > 0 && !opDims) || nd < 0 || nop < 0 || nsrc != ndst || (nsrc > 0 && (!src || !dst))) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nd > kMaxDims) {
            throw std::runtime_error("Too many dimensions for the CUDA provider");
        }
        check_dimensions(sD, sS, nd, nsrc);
        std::vector<char> flag((size_t) nd, 0);
        for (int i = 0; i < nop; i++) {
            if (!(opDims[i] >= 0 && opDims[i] < nd)) {
                throw std::runtime_error("Invalid dimension");
            }
            flag[(size_t) opDims[i]] = 1;
        }
        if (nsrc == 0) {
            return;
        }
        // both dimensions of a dense 2-D array: one pass over the data instead of one per dimension (array_scan2d.cu)
        if (type == RD_SUM && nd == 2 && flag[0] && flag[1] && sD[0] > 1 && sD[1] > 1) {
            const bool row_major = sS[1] == 1 && sS[0] == sD[1], col_major = sS[0] == 1 && sS[1] == sD[0];
            const int64_t rows = row_major ? sD[0] : sD[1], cols = row_major ? sD[1] : sD[0];
            if ((row_major || col_major) && scan2d_applicable(rows, cols)) {
                scan2d_sum(src, cols, dst, cols, rows, cols, false, s);
                return;
            }
        }
        std::vector<int64_t> d = widen(sD, nd), st = widen(sS, nd);
        const double *cur = src;
        bool any = false;
        for (int dim = 0; dim < nd; dim++) {  // ascending, one dimension after another (DimensionOps.java:411-435)
            if (!flag[(size_t) dim]) {
                continue;
            }
            Lines L = make_lines(d, st, st, dim);
            if (type == RD_SUM) {
                scan_dim<RD_SUM>(cur, dst, L, s);
            } else {
                scan_dim<RD_PROD>(cur, dst, L, s);
            }
            cur = dst;
            any = true;
        }
        if (!any && src != dst) {
            SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) nsrc * 8, cudaMemcpyDeviceToDevice, s));
        }
    });
}

int sstc_riop_dev(int type, double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, int32_t *dst,
        int64_t ndst, int dim, void *stream) {
    return guarded([&] {
        cudaStream_t s = as_stream(stream);
        if (type < RI_MAX || type > RI_SORT) {
            throw std::runtime_error("Operation type not recognized");
        }
        if (!sD || !sS || nd < 0 || nsrc != ndst || (nsrc > 0 && (!src || !dst))) {
            throw std::runtime_error("Invalid arguments");
        }
        if (nd > kMaxDims) {
            throw std::runtime_error("Too many dimensions for the CUDA provider");
        }
        check_dimensions(sD, sS, nd, nsrc);
        if (!((dim >= 0 && dim < nd) || dim == -1)) {
            throw std::runtime_error("Invalid dimension");
        }
        if (nsrc == 0) {
            return;
        }
        if (dim == -1) {
            const int grid = stream_grid(nsrc);
            if (type == RI_SORT) {
                if (nsrc > INT_MAX) {
                    throw std::runtime_error("Invalid arguments");
                }
                std::vector<int64_t> d1(1, nsrc), s1(1, 1);
                Lines L = make_lines(d1, s1, s1, 0);
                sort_dim(src, dst, L, s);
                return;
            }
            double *ext = nullptr;
            if (type == RI_MAX || type == RI_MIN) {
                ext = static_cast<double *>(t_tmp[2].get(16));
                if (sstc_raop_dev(type == RI_MAX ? RA_MAX : RA_MIN, src, nsrc, ext, stream) != 0) {
                    throw std::runtime_error(sstc_last_error());
                }
            }
            switch (type) {
            case RI_MAX: index_mask_kernel<RI_MAX><<<grid, kThreads, 0, s>>>(src, dst, nsrc, ext); break;
            case RI_MIN: index_mask_kernel<RI_MIN><<<grid, kThreads, 0, s>>>(src, dst, nsrc, ext); break;
            case RI_ZERO: index_mask_kernel<RI_ZERO><<<grid, kThreads, 0, s>>>(src, dst, nsrc, ext); break;
            case RI_GZERO: index_mask_kernel<RI_GZERO><<<grid, kThreads, 0, s>>>(src, dst, nsrc, ext); break;
            default: index_mask_kernel<RI_LZERO><<<grid, kThreads, 0, s>>>(src, dst, nsrc, ext); break;
            }
            check_launch("index_mask_kernel");
            return;
        }
        std::vector<int64_t> d = widen(sD, nd), st = widen(sS, nd);
        Lines L = make_lines(d, st, st, dim);
        switch (type) {
        case RI_MAX: index_dim<RI_MAX>(src, dst, L, s); break;
        case RI_MIN: index_dim<RI_MIN>(src, dst, L, s); break;
        case RI_ZERO: index_dim<RI_ZERO>(src, dst, L, s); break;
        case RI_GZERO: index_dim<RI_GZERO>(src, dst, L, s); break;
        case RI_LZERO: index_dim<RI_LZERO>(src, dst, L, s); break;
        default: sort_dim(src, dst, L, s); break;
        }
    });
}

}  // extern "C"

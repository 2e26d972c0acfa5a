// fft_internal.h -- what the FFT translation units share: plans and their steps (fft_plan.cu), the host-tier pipeline
// (fft_host.cu) and the slab-decomposed multi-GPU transform (fft_slab.cu). Kernels stay in fft_kernels.cuh, which only
// fft_plan.cu includes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "sstc_internal.h"

namespace sstc {

enum FftPassMode : int { PASS_C2C = 0, PASS_R2C = 1, PASS_C2R = 2 };  // == FftMode in fft_kernels.cuh

struct Step {
    int axis_len;
    int64_t outer, inner;
    int mode;
    int src;  // 0 = caller's in, 1 = caller's out, 2 = workspace
    int dst;  // 1 = caller's out, 2 = workspace
    double scale;
    int64_t layout[4];  // {in_plane, in_kstride, out_plane, out_kstride}; all 0 = the natural layout
};

}  // namespace sstc

struct sstc_fft_plan_s {
    int kind;
    std::vector<int32_t> dims;
    int64_t in_len, out_len;  // doubles
    int inverse;
    std::vector<sstc::Step> steps;
    void *workspace = nullptr;
    int64_t workspace_bytes = 0;
    int transposed_passes = -1;  // hint: -1 auto, 0 never, 1 whenever legal
    int alternate_order = 0;     // hint: odd passes process their tiles in descending order (measured: no gain)
};

namespace sstc {

void transform_lengths(int kind, int rank, const int32_t *dims, int64_t &in_len, int64_t &out_len);
void build_plan(sstc_fft_plan_s &p);
void execute_plan(sstc_fft_plan_s &p, const double *d_in, double *d_out, cudaStream_t stream);
sstc_fft_plan_s *new_plan(int kind, int rank, const int32_t *dims);
void delete_plan(sstc_fft_plan_s *p);

// true when one tile kernel transforms an axis of this length (anything else runs four-step / Bluestein, which need
// the natural layout)
bool tile_kernels_hold(int L);

// One pass over axis L of (outer, L, inner); layout = {in_plane, in_kstride, out_plane, out_kstride} or nullptr for the
// natural one. `mode` is a FftPassMode (R2C / C2R: inner == 1, outer counts lines).
void fft_axis_pass(const void *in, void *out, int64_t outer, int L, int64_t inner, int mode, int inverse, double scale,
        cudaStream_t stream, const int64_t *layout = nullptr);

// The exchange passes of the slab transform (documented at sstc_fft_lines_scatter_dev / sstc_fft_planes_scatter_dev).
// bulk = 1: each transformed line leaves shared memory as ONE bulk asynchronous copy (cp.async.bulk shared -> global)
// instead of per-thread 16-byte stores.
void fft_lines_scatter(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes, int32_t plane_lines,
        int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas, int bulk, cudaStream_t stream);
void fft_planes_scatter(const double *d_in, void *const *d_out_blocks, int nblocks, int32_t block_planes, int32_t plane_lines,
        int32_t dst_plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas,
        int bulk, cudaStream_t stream);

// Host-side helpers of the host tier (fft_host.cu): copies spread over the staging threads, and whether a host pointer is
// page-locked (cudaMallocHost / cudaHostRegister) so that the DMA engines can read it directly.
void host_copy(void *dst, const void *src, size_t bytes);
void host_copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows);
bool host_is_pinned(const void *p);

}  // namespace sstc

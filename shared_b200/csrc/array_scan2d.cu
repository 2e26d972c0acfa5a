// array_scan2d.cu -- single-pass 2-D inclusive sum scan (summed-area table) of a dense row-major array.
//
// Serves rdOp(RD_SUM) over both dimensions of a 2-D array (ArrayKernel.java:438, DimensionOps.java:411-435 applies the
// scan dimension by dimension) and ImageKernel.createIntegralImage (NativeImageKernel.cpp:32-118: copy the source to
// offset (1, 1) of the destination, then scan the WHOLE destination along dimension 0 and along dimension 1). Scanning
// dimension after dimension costs one HBM round trip per dimension (32 B per element for two); here every element is
// read once and written once (16 B): the array is cut into 64 x 64 tiles, a tile's row sums and column sums travel to
// the tiles right of and below it through small descriptors with DECOUPLED LOOK-BACK (a tile publishes its own
// aggregate as soon as it has it, and adds up its predecessors' aggregates until it meets one that already knows its
// inclusive prefix), so no tile waits for more than the latency of a few descriptor reads.
//
//   Y = row scan of the input:     Y[i][j] = sum_{b <= j} I[i][b]      carry of a tile: per ROW, the sum over the tiles left
//   S = column scan of Y:          S[i][j] = sum_{a <= i} Y[a][j]      carry of a tile: per COLUMN, S of the row above it
//
// Summation order differs from the reference's (a tile's partial sums are combined tree-wise), which is within the
// scans' floating-point bar (relative L2 <= 1e-12 log2 N) and exact whenever the partial sums are exactly representable
// (integer-valued inputs), as the parity tests check.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "array_common.cuh"
#include "scan2d_core.cuh"

namespace sstc {

constexpr int kScanThreads = 256;   // 8 warps: warp w owns rows 8w..8w+7 of the tile for the row scan

struct Scan2dArgs {
    const double *src;
    double *dst;
    int64_t rows, cols, spitch, dpitch;
    int tiles_r, tiles_c;
    ScanDesc *row_desc, *col_desc;  // [tiles_r * tiles_c], indexed by tile position
    unsigned int *ticket;
};

// the descriptor words as the kernel touches them (scan2d_core.cuh's memory policy)
struct DeviceScanMem {
    static __device__ __forceinline__ void load_pair(const ScanPair *p, unsigned long long &agg, unsigned long long &pre) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(agg), "=l"(pre) : "l"(p) : "memory");
    }
    static __device__ __forceinline__ void store(double *p, unsigned long long bits) {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(bits) : "memory");
    }
};

__device__ __forceinline__ void publish(double *p, double v) { publish<DeviceScanMem>(p, v); }

__device__ __forceinline__ double look_back(const ScanDesc *desc, int64_t pos, int64_t step, int count, int e) {
    return look_back<DeviceScanMem>(desc, pos, step, count, e);
}

// BORDER: the input is the destination with its interior replaced by the source shifted by (1, 1) (createIntegralImage).
template <bool BORDER>
__global__ void __launch_bounds__(kScanThreads) scan2d_kernel(const Scan2dArgs a) {
    __shared__ double tile[kT][kT + 1];
    __shared__ double row_carry[kT], col_carry[kT], col_part[4][kT];
    __shared__ int s_tr[2], s_tc[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ntiles = (int64_t) a.tiles_r * a.tiles_c;

    // Tiles are handed out along anti-diagonals: both predecessors of a tile lie on the previous diagonal, so every tile
    // a CTA may wait for was taken earlier by a CTA that is running. The NEXT ticket is drawn while the current tile is
    // being loaded, so that the atomic's round trip is off the critical path.
    auto draw = [&](int slot) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        if ((int64_t) t >= ntiles) {
            s_tr[slot] = -1;
        } else {
            ticket_to_tile(t, a.tiles_r, a.tiles_c, s_tr[slot], s_tc[slot]);
        }
    };
    if (tid == 0) {
        draw(0);
    }
    __syncthreads();
    for (int it = 0;; it++) {
        const int tr = s_tr[it & 1], tc = s_tc[it & 1];
        if (tr < 0) {
            return;
        }
        if (tid == kScanThreads - 1) {
            draw((it + 1) & 1);
        }
        const int64_t pos = (int64_t) tr * a.tiles_c + tc;
        const int64_t r0 = (int64_t) tr * kT, c0 = (int64_t) tc * kT;

        // ---- load + row scan: warp w, rows 8w..8w+7; a lane owns columns lane and lane + 32 of a row ----------------
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int r = warp * 8 + k;
            const int64_t gi = r0 + r;
            double x0 = 0.0, x1 = 0.0;
            if (gi < a.rows) {
                const int64_t g0 = c0 + lane, g1 = c0 + lane + 32;
                if (BORDER) {
                    if (g0 < a.cols) {
                        x0 = (gi == 0 || g0 == 0) ? a.dst[gi * a.dpitch + g0] : a.src[(gi - 1) * a.spitch + (g0 - 1)];
                    }
                    if (g1 < a.cols) {
                        x1 = gi == 0 ? a.dst[gi * a.dpitch + g1] : a.src[(gi - 1) * a.spitch + (g1 - 1)];
                    }
                } else {
                    if (g0 < a.cols) {
                        x0 = a.src[gi * a.spitch + g0];
                    }
                    if (g1 < a.cols) {
                        x1 = a.src[gi * a.spitch + g1];
                    }
                }
            }
            // inclusive scans of the two 32-wide halves
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double y0 = __shfl_up_sync(0xffffffffu, x0, d), y1 = __shfl_up_sync(0xffffffffu, x1, d);
                if (lane >= d) {
                    x0 += y0;
                    x1 += y1;
                }
            }
            x1 += __shfl_sync(0xffffffffu, x0, 31);
            tile[r][lane] = x0;
            tile[r][lane + 32] = x1;
            if (lane == 31) {
                publish(&a.row_desc[pos].ap[r].x, x1);  // the row's sum over this tile
                row_carry[r] = x1;                     // parked here until the carry is known
            }
        }
        __syncthreads();
        // ---- row carry: the sums of the tiles on the left ---------------------------------------------------------------
        if (tid < kT) {
            const double c = look_back(a.row_desc, pos, 1, tc, tid);
            publish(&a.row_desc[pos].ap[tid].y, c + row_carry[tid]);
            row_carry[tid] = c;
        }
        __syncthreads();
        // ---- Y = row scan + carry; column sums of Y: thread (q, c) sums rows 16q..16q+15 of column c --------------------
        const int c = tid & (kT - 1), q = tid >> 6;
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int r = q * 16 + k;
            const double y = tile[r][c] + row_carry[r];
            tile[r][c] = y;
            part += y;
        }
        col_part[q][c] = part;
        __syncthreads();
        // ---- column carry: S of the row just above the tile ---------------------------------------------------------------
        if (tid < kT) {
            const double own = col_part[0][tid] + col_part[1][tid] + col_part[2][tid] + col_part[3][tid];
            publish(&a.col_desc[pos].ap[tid].x, own);
            const double cc = look_back(a.col_desc, pos, a.tiles_c, tr, tid);
            publish(&a.col_desc[pos].ap[tid].y, cc + own);
            col_carry[tid] = cc;
        }
        __syncthreads();
        // ---- column scan and store: thread (q, c) continues from the parts above it -----------------------------------------
        double run = col_carry[c];
        for (int k = 0; k < q; k++) {
            run += col_part[k][c];
        }
        const int64_t gj = c0 + c;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int r = q * 16 + k;
            run += tile[r][c];
            const int64_t gi = r0 + r;
            if (gi < a.rows && gj < a.cols) {
                a.dst[gi * a.dpitch + gj] = run;
            }
        }
        __syncthreads();  // the tile buffers are reused by the next ticket
    }
}

static thread_local Scratch t_scan_ws;

bool scan2d_applicable(int64_t rows, int64_t cols) {
    // below this the per-dimension kernels are launch-bound either way; above, descriptors must fit 32-bit tile counts
    return rows >= 2 && cols >= 2 && rows * cols >= (1 << 16) && rows <= (1 << 24) && cols <= (1 << 24) &&
            ((rows + kT - 1) / kT) * ((cols + kT - 1) / kT) < (1LL << 31);
}

// dst = 2-D inclusive sum scan of the (rows x cols) input; border = createIntegralImage's virtual input (see above).
// In place (src == dst, equal pitches) is allowed: a tile reads its input before it writes its output and nothing
// else reads either.
void scan2d_sum(const double *src, int64_t spitch, double *dst, int64_t dpitch, int64_t rows, int64_t cols, bool border,
        cudaStream_t stream) {
    Scan2dArgs a;
    a.src = src;
    a.dst = dst;
    a.rows = rows;
    a.cols = cols;
    a.spitch = spitch;
    a.dpitch = dpitch;
    a.tiles_r = (int) ((rows + kT - 1) / kT);
    a.tiles_c = (int) ((cols + kT - 1) / kT);
    const int64_t nt = (int64_t) a.tiles_r * a.tiles_c;
    const size_t desc_bytes = (size_t) nt * sizeof(ScanDesc);
    char *ws = static_cast<char *>(t_scan_ws.get((int64_t) (2 * desc_bytes + 256)));
    a.row_desc = reinterpret_cast<ScanDesc *>(ws);
    a.col_desc = reinterpret_cast<ScanDesc *>(ws + desc_bytes);
    a.ticket = reinterpret_cast<unsigned int *>(ws + 2 * desc_bytes);
    SSTC_CUDA(cudaMemsetAsync(ws, 0xFF, 2 * desc_bytes, stream));  // every descriptor word = "not ready"
    SSTC_CUDA(cudaMemsetAsync(a.ticket, 0, sizeof(unsigned int), stream));
    static const int per_sm = getenv("SSTC_SCAN2D_CTAS_PER_SM") ? std::max(1, atoi(getenv("SSTC_SCAN2D_CTAS_PER_SM"))) : 6;
    const int grid = (int) std::min<int64_t>(nt, (int64_t) sm_count() * per_sm);
    if (border) {
        scan2d_kernel<true><<<grid, kScanThreads, 0, stream>>>(a);
    } else {
        scan2d_kernel<false><<<grid, kScanThreads, 0, stream>>>(a);
    }
    check_launch("scan2d_kernel");
}

}  // namespace sstc

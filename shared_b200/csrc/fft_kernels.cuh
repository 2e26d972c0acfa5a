// fft_kernels.cuh -- sm_100a device code for one axis pass of an n-D complex128 FFT.
//
// A row-major complex array is viewed as (outer, L, inner); one pass transforms every length-L line
// (o, :, i). The reference does the same work one line at a time on the CPU
// (src/org/shared/fft/FftOps.java:65-84 per-dimension loop, :173-248 butterflies); here a CTA owns a tile of
// C lines and runs a Stockham autosort FFT on it:
//
//   * each thread keeps 8 points of one line in registers; thread t of a line owns points t + m*(L/8),
//     m = 0..7, in EVERY stage, so the first stage is fed straight from HBM and the last stage stores
//     straight to HBM -- both as coalesced 16-byte (double2) accesses;
//   * stages are radix 8 (then one radix 4 or 2 if log2 L is not a multiple of 3), computed in registers;
//     between stages the tile is exchanged through shared memory (one write + one read of 16 B / point);
//   * CONTIG tiles (inner == 1) hold C whole lines, STRIDED tiles (inner > 1) hold C adjacent columns of
//     the (L x inner) plane so that every row segment of the tile is C*16 contiguous bytes in HBM;
//   * the inverse transform reuses the forward butterflies through the swap identity
//     ifft(x) = swap(fft(swap(x))), swap(a+ib) = b+ia, applied on load and store;
//   * r2c / c2r (last axis only) are load/store variants of the same kernel: r2c promotes a real line and
//     keeps bins [0, L/2]; c2r rebuilds the Hermitian line x[L-k] = conj(x[k]) on load and stores Re.
//
// HBM traffic per pass is the algorithmic minimum for an axis-separable schedule: 16 B read + 16 B written
// per point.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sstc {

enum FftMode : int { MODE_C2C = 0, MODE_R2C = 1, MODE_C2R = 2 };

struct FftPassArgs {
    const double2 *in;   // c2c / c2r: complex input; r2c: reinterpret as const double*
    double2 *out;        // c2c / r2c: complex output; c2r: reinterpret as double*
    const double2 *tw;   // generic kernel: tw[m] = exp(-2 pi i m / L), m in [0, L)
                         // pow2 kernel: stage-major table, see pow2_twiddle_count()
    const double2 *tw2;  // real-transform kernel: tw2[k] = exp(-2 pi i k / N), k in [0, N/2)
    int64_t outer;       // number of (L x inner) planes
    int64_t inner;       // stride between consecutive points of a line, in elements
    int64_t in_plane;    // elements between consecutive planes of the input (== L*inner unless sub-viewed)
    int64_t out_plane;
    int64_t in_kstride;  // STRIDED: elements between consecutive points of a line (== inner unless the pass
    int64_t out_kstride; //          also transposes, e.g. reads (o,k,i) and writes (k,o,i))
    double scale;        // applied on store
    int mode;            // FftMode
    int inverse;         // 0: e^{-i..}, 1: e^{+i..}
    // Blocked addressing (STRIDED c2c only; ksplit == 0 = off). The axis is cut into blocks of `ksplit`
    // points; point k of plane o, column i lives at blk[k / ksplit] + (o*ksplit + k % ksplit)*pitch + i.
    // On the store side, with blk[] pointing into the peers' receive buffers, this IS the slab transpose of
    // the multi-GPU transform (pack, send and unpack fused into the pass's store); on the load side it reads
    // an axis that an all-to-all delivered as one block per source rank.
    int ksplit;
    int in_ksplit;
    int64_t blk_pitch;
    int64_t in_blk_pitch;
    double2 *blk[8];
    const double2 *in_blk[8];
    // Line scatter (CONTIG c2c only; ls_group == 0 = off): the input is (planes, ls_plane_lines, L) and the
    // lines of a plane are cut into ls_nblk blocks of ls_block_lines; this launch transforms lines
    // [ls_first, ls_first + ls_group) of every block of every plane and stores line yl of block q of plane xl
    // at blk[q] + (xl*ls_block_lines + yl)*L -- whole 16*L-byte lines per destination, which is what lets
    // the exchange pass of the slab transform run in groups that the next pass can start consuming.
    // ls_sub > 0 selects the mirrored form used by the inverse slab transform: the input is (nblk, ls_block_lines
    // planes, ls_sub lines, L); this launch transforms every line of planes [ls_first, ls_first + ls_group) of every
    // block and stores line yl of plane xl of block q at blk[q] + (xl*ls_plane_lines + yl)*L (ls_plane_lines = the
    // destination's lines per plane).
    int ls_group, ls_nblk, ls_first, ls_block_lines, ls_plane_lines, ls_sub;
    // ls_bulk: the transformed line is put back into shared memory and leaves as ONE bulk asynchronous copy
    // (cp.async.bulk shared::cta -> global, SASS UBLKCP) issued by one thread per line, instead of 16-byte stores from
    // every thread: the copy engine of the SM streams the 16*L bytes to the peer and the LSU is free for the next tile.
    int ls_bulk;
    // Fused convolution (EXT kernels and the any-length kernel; all off when mul == nullptr and crop_cols == 0):
    //   mul        every point loaded by a c2c (natural layout) or c2r pass is first multiplied by mul[same index]
    //              -- the eMul of ConvolutionCache.convolve folded into the first inverse pass;
    //   crop_cols  a c2r pass stores only the first crop_cols reals of a line, at line pitch crop_cols, and its
    //              lines are numbered in the CROPPED leading dims crop_out[] (row-major) and mapped back to the
    //              full leading dims crop_in[] on load -- the subarray of convolve folded into the last pass.
    const double2 *mul;
    int crop_cols, crop_nd;
    int32_t crop_in[4], crop_out[4];
    int64_t ntiles;      // tiles in this launch (set by the launcher); the grid may be smaller
    int64_t grid_limit;  // 0 = one CTA per tile
    int reverse;         // process the tiles in descending order
    // STRIDED hot path only: every thread of a tile also issues one prefetch.global.L2 for a row segment of the tile
    // `prefetch_ahead` tiles further on (0 = off), so that the page walks and the DRAM reads of a tile with a huge row
    // stride (the outermost pass: every row in its own 2 MiB page) are under way before its CTA starts.
    int prefetch_ahead;
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 w) {
    return make_double2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
// a * (-i)
__device__ __forceinline__ double2 cmul_mi(double2 a) { return make_double2(a.y, -a.x); }

// Line number in the cropped leading dims -> line number in the full leading dims (see FftPassArgs::crop_cols).
__device__ __forceinline__ int64_t crop_source_line(const FftPassArgs &a, int64_t line) {
    int64_t src = 0, mult = 1;
    for (int d = a.crop_nd - 1; d >= 0; d--) {
        const int64_t q = line / a.crop_out[d];
        src += (line - q * a.crop_out[d]) * mult;
        mult *= a.crop_in[d];
        line = q;
    }
    return a.crop_nd > 0 ? src : line;
}

// In-register DFTs over v[0], v[S], v[2S], ... (forward sign), outputs in natural order at the same slots.
template <int S>
__device__ __forceinline__ void dft2(double2 *v) {
    double2 a = v[0], b = v[S];
    v[0] = cadd(a, b);
    v[S] = csub(a, b);
}

template <int S>
__device__ __forceinline__ void dft4(double2 *v) {
    double2 c0 = cadd(v[0], v[2 * S]), c1 = csub(v[0], v[2 * S]);
    double2 c2 = cadd(v[S], v[3 * S]), c3 = cmul_mi(csub(v[S], v[3 * S]));
    v[0] = cadd(c0, c2);
    v[S] = cadd(c1, c3);
    v[2 * S] = csub(c0, c2);
    v[3 * S] = csub(c1, c3);
}

__device__ __forceinline__ void dft8(double2 *v) {
    const double h = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), d5 = csub(v[1], v[5]);
    double2 a2 = cadd(v[2], v[6]), a6 = cmul_mi(csub(v[2], v[6]));
    double2 a3 = cadd(v[3], v[7]), d7 = csub(v[3], v[7]);
    // d5 * (1 - i)/sqrt2 ; d7 * (-1 - i)/sqrt2
    double2 a5 = make_double2((d5.x + d5.y) * h, (d5.y - d5.x) * h);
    double2 a7 = make_double2((d7.y - d7.x) * h, -(d7.x + d7.y) * h);
    // even outputs: DFT4(a0,a1,a2,a3); odd outputs: DFT4(a4,a5,a6,a7)
    double2 c0 = cadd(a0, a2), c1 = csub(a0, a2), c2 = cadd(a1, a3), c3 = cmul_mi(csub(a1, a3));
    v[0] = cadd(c0, c2);
    v[2] = cadd(c1, c3);
    v[4] = csub(c0, c2);
    v[6] = csub(c1, c3);
    double2 e0 = cadd(a4, a6), e1 = csub(a4, a6), e2 = cadd(a5, a7), e3 = cmul_mi(csub(a5, a7));
    v[1] = cadd(e0, e2);
    v[3] = cadd(e1, e3);
    v[5] = csub(e0, e2);
    v[7] = csub(e1, e3);
}

template <int R, int B>
__device__ __forceinline__ void dftR(double2 *v) {
    if (R == 8) {
        dft8(v);
    } else if (R == 4) {
        dft4<B>(v);
    } else {
        dft2<B>(v);
    }
}

__host__ __device__ constexpr int fft_pad(int pos) { return pos + (pos >> 3); }

// Radix of the Stockham stage that starts with `ns` already-combined points (radix 8 while it fits).
__host__ __device__ constexpr int pow2_stage_radix(int L, int ns) { return (L / ns) >= 8 ? 8 : (L / ns); }

// The pow2 kernel's twiddle table is stage-major so that a warp's lanes (consecutive k) read consecutive
// 16-byte entries: for every stage after the first, entries [(r-1)*NS + k] = exp(-2 pi i r k / (NS*R)),
// r = 1..R-1, k = 0..NS-1, stages concatenated in execution order.
__host__ __device__ constexpr int pow2_twiddle_count(int L) {
    int total = 0;
    for (int ns = pow2_stage_radix(L, 1); ns < L; ns *= pow2_stage_radix(L, ns)) {
        total += (pow2_stage_radix(L, ns) - 1) * ns;
    }
    return total;
}

// Tile geometry shared by host (launch configuration) and device.
template <int L, int C, bool STRIDED>
struct Pow2Tile {
    static constexpr int T = L / 8;                 // threads per line
    static constexpr int THREADS = C * T;
    static constexpr bool PADDED = !STRIDED || (C < 8);
    static constexpr int LINE = PADDED ? fft_pad(L) : L;  // smem elements per line
    static constexpr int SMEM_BYTES = (L > 8) ? LINE * C * 16 : 0;
    // keep 1024 threads resident per SM (<= 64 registers per thread): the passes are latency-bound otherwise
    static constexpr int MIN_CTAS = (1024 / THREADS) < 1 ? 1 : (1024 / THREADS) > 8 ? 8 : (1024 / THREADS);
};

// One Stockham stage with Ns = NS already-combined points, followed (if not last) by the smem exchange and
// the remaining stages.
template <int L, int C, bool STRIDED, int NS>
struct Pow2Stages {
    static constexpr int REM = L / NS;
    static constexpr int R = REM >= 8 ? 8 : REM;
    static constexpr int B = 8 / R;
    static constexpr int T = L / 8;
    using Tile = Pow2Tile<L, C, STRIDED>;

    __device__ __forceinline__ static int sidx(int pos, int c) {
        int p = Tile::PADDED ? fft_pad(pos) : pos;
        return STRIDED ? p * C + c : c * Tile::LINE + p;
    }

    __device__ __forceinline__ static void run(double2 (&v)[8], int t, int c, double2 *sm,
            const double2 *__restrict__ tw) {
        if (NS > 1) {
            if (R == 8 && NS >= 256) {
                // The last stage of a long line: its 7 * NS table (57 KB at L = 4096) does not stay in L1 beside the tile
                // traffic, and seven 16-byte twiddle loads per thread are nearly as much L2 -> SM traffic as the data.
                // Three loads (w^k, w^2k, w^4k) and four complex products give the other four powers.
                const int k = t & (NS - 1);
                const double2 w1 = __ldg(&tw[k]), w2 = __ldg(&tw[NS + k]), w4 = __ldg(&tw[3 * NS + k]);
                const double2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
                const double2 w7 = cmul(w4, w3);
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                v[5] = cmul(v[5], w5);
                v[6] = cmul(v[6], w6);
                v[7] = cmul(v[7], w7);
            } else {
#pragma unroll
                for (int b = 0; b < B; b++) {
                    int k = (t + b * T) & (NS - 1);
#pragma unroll
                    for (int r = 1; r < R; r++) {
                        double2 w = __ldg(&tw[(r - 1) * NS + k]);  // w_{NS*R}^{r*k}; lanes read consecutive k
                        v[b + r * B] = cmul(v[b + r * B], w);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < B; b++) {
            dftR<R, B>(&v[b]);
        }
        if (NS * R < L) {
            if (NS > 1) {
                __syncthreads();  // everyone has finished reading the previous exchange
            }
#pragma unroll
            for (int b = 0; b < B; b++) {
                int j = t + b * T;
                int k = j & (NS - 1);
                int j0 = (j - k) * R + k;
#pragma unroll
                for (int q = 0; q < R; q++) {
                    sm[sidx(j0 + q * NS, c)] = v[b + q * B];
                }
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 8; m++) {
                v[m] = sm[sidx(t + m * T, c)];
            }
            Pow2Stages<L, C, STRIDED, (NS * R < L ? NS * R : L)>::run(v, t, c, sm,
                    tw + (NS > 1 ? (R - 1) * NS : 0));
        }
    }
};

template <int L, int C, bool STRIDED>
struct Pow2Stages<L, C, STRIDED, L> {
    __device__ __forceinline__ static void run(double2 (&)[8], int, int, double2 *, const double2 *) {}
};

// One tile (C lines of length L) of one pass; `blk` is the tile number. EXT = false is the lean hot path
// (natural layout); EXT = true adds explicit point strides, blocked addressing and the line scatter, which
// cost registers the hot path cannot spare at 1024 resident threads per SM.
template <int L, int C, bool STRIDED, bool EXT, bool PF = false>
__device__ __forceinline__ void fft_pow2_tile(const FftPassArgs &a, const int64_t blk, double2 *sm) {
    using Tile = Pow2Tile<L, C, STRIDED>;
    constexpr int T = Tile::T;

    const int tid = threadIdx.x;
    int c, t;
    if (STRIDED) {
        c = tid % C;
        t = tid / C;
    } else {
        t = tid % T;
        c = tid / T;
    }

    // Locate the tile. STRIDED: blk -> (plane o, column tile); CONTIG: blk -> C lines.
    bool active;
    int64_t in_base, out_base, in_ks, out_ks;
    int64_t line = 0;  // CONTIG line number (for r2c / c2r addressing)
    int64_t o = 0, i = 0;
    double2 *out_line = nullptr;  // CONTIG line scatter destination
    if (STRIDED) {
        const int64_t tiles_per_plane = (a.inner + C - 1) / C;
        // 32-bit division whenever both fit (always, short of 2^32 tiles): the 64-bit one is ~100 dependent
        // instructions in front of the first load of every CTA
        o = ((blk | tiles_per_plane) >> 32) == 0 ? (int64_t) ((uint32_t) blk / (uint32_t) tiles_per_plane)
                                                  : blk / tiles_per_plane;
        i = (blk - o * tiles_per_plane) * C + c;
        active = i < a.inner;
        in_base = o * a.in_plane + i;
        out_base = o * a.out_plane + i;
        in_ks = EXT ? a.in_kstride : a.inner;
        out_ks = EXT ? a.out_kstride : a.inner;
    } else {
        line = blk * C + c;
        active = line < a.outer;
        in_base = line * a.in_plane;
        out_base = line * a.out_plane;
        in_ks = out_ks = 1;
        if (EXT && a.ls_group > 0) {
            // line scatter: this launch covers, for every plane xl and every destination block q, the lines
            // [ls_first, ls_first + ls_group) of that block; a whole line goes to one destination.
            if (a.ls_sub > 0) {
                const int64_t yl = line % a.ls_sub, t1 = line / a.ls_sub;
                const int q = (int) (t1 % a.ls_nblk);
                const int64_t xl = a.ls_first + t1 / a.ls_nblk;
                in_base = (((int64_t) q * a.ls_block_lines + xl) * a.ls_sub + yl) * L;
                out_line = a.blk[q] + (xl * a.ls_plane_lines + yl) * L;
            } else {
                const int64_t yq = line % a.ls_group, t1 = line / a.ls_group;
                const int q = (int) (t1 % a.ls_nblk);
                const int64_t xl = t1 / a.ls_nblk;
                const int64_t yl = a.ls_first + yq;
                in_base = (xl * a.ls_plane_lines + (int64_t) q * a.ls_block_lines + yl) * L;
                out_line = a.blk[q] + (xl * a.ls_block_lines + yl) * L;
            }
        }
    }

    double2 v[8];
    if (active) {
        if (EXT && STRIDED && a.in_ksplit > 0) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int p = t + m * T;
                const int q = p / a.in_ksplit;
                const int pk = p - q * a.in_ksplit;
                v[m] = a.in_blk[q][(o * a.in_ksplit + pk) * a.in_blk_pitch + i];
            }
        } else if (a.mode == MODE_C2C) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                v[m] = a.in[in_base + (int64_t) (t + m * T) * in_ks];
            }
            if (PF && !EXT && STRIDED) {
                const int64_t fb = blk + a.prefetch_ahead;
                if (fb < (int64_t) gridDim.x) {
                    const int64_t tiles_per_plane = (a.inner + C - 1) / C;
                    const int64_t fo = fb / tiles_per_plane, fi = (fb - fo * tiles_per_plane) * C;
                    // thread tid -> row tid of the future tile (THREADS == L whenever C == 8)
                    for (int r = tid; r < L; r += Tile::THREADS) {
                        const double2 *p = a.in + fo * a.in_plane + (int64_t) r * a.inner + fi;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                    }
                }
            }
            if (EXT && a.mul) {
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    v[m] = cmul(v[m], a.mul[in_base + (int64_t) (t + m * T) * in_ks]);
                }
            }
        } else if (a.mode == MODE_R2C) {
            const double *rin = reinterpret_cast<const double *>(a.in) + line * L;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                v[m] = make_double2(rin[t + m * T], 0.0);
            }
        } else {  // MODE_C2R: half-complex line of L/2+1 bins
            const int64_t hoff = (EXT && a.crop_cols > 0 ? crop_source_line(a, line) : line) * (L / 2 + 1);
            const double2 *hin = a.in + hoff;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int p = t + m * T;
                const int q = p <= L / 2 ? p : L - p;
                double2 z = hin[q];
                if (EXT && a.mul) {
                    z = cmul(z, a.mul[hoff + q]);
                }
                v[m] = p <= L / 2 ? z : make_double2(z.x, -z.y);
            }
        }
        if (a.inverse) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                v[m] = make_double2(v[m].y, v[m].x);
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            v[m] = make_double2(0.0, 0.0);
        }
    }

    Pow2Stages<L, C, STRIDED, 1>::run(v, t, c, sm, a.tw);

    const double s = a.scale;
    if (EXT && !STRIDED && L >= 16 && a.ls_group > 0 && a.ls_bulk) {
        // every thread takes part (barriers), inactive lines just do not issue a copy
        __syncthreads();  // the last exchange has been read by everyone
        double2 *ln = sm + c * L;  // dense line (the padded exchange layout is not contiguous)
#pragma unroll
        for (int m = 0; m < 8; m++) {
            ln[t + m * T] = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
        __syncthreads();
        if (t == 0 && active) {
            const uint32_t src = (uint32_t) __cvta_generic_to_shared(ln);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_line), "r"(src),
                         "r"((uint32_t) (L * 16))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        return;
    }
    if (!active) {
        return;
    }
    if (EXT && STRIDED && a.ksplit > 0) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            double2 z = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
            const int p = t + m * T;
            const int q = p / a.ksplit;
            const int pk = p - q * a.ksplit;
            a.blk[q][(o * a.ksplit + pk) * a.blk_pitch + i] = z;
        }
    } else if (EXT && !STRIDED && a.ls_group > 0) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            double2 z = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
            out_line[t + m * T] = z;
        }
    } else if (a.mode == MODE_C2C) {
        if (!a.inverse && s == 1.0) {  // the unscaled forward transform: nothing to multiply
#pragma unroll
            for (int m = 0; m < 8; m++) {
                a.out[out_base + (int64_t) (t + m * T) * out_ks] = v[m];
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                double2 z = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
                a.out[out_base + (int64_t) (t + m * T) * out_ks] = z;
            }
        }
    } else if (a.mode == MODE_R2C) {
        double2 *hout = a.out + line * (L / 2 + 1);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int p = t + m * T;
            if (p <= L / 2) {
                hout[p] = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
            }
        }
    } else if (EXT && a.crop_cols > 0) {
        double *rout = reinterpret_cast<double *>(a.out) + line * a.crop_cols;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            if (t + m * T < a.crop_cols) {
                rout[t + m * T] = (a.inverse ? v[m].y : v[m].x) * s;
            }
        }
    } else {
        double *rout = reinterpret_cast<double *>(a.out) + line * L;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            rout[t + m * T] = (a.inverse ? v[m].y : v[m].x) * s;
        }
    }
}

// Grid = one CTA per tile normally; a smaller grid makes the CTAs loop over the tiles, which is how the
// NVLink-bound exchange pass is confined to a fraction of every SM so that the next pass can run beside it.
template <int L, int C, bool STRIDED>
__global__ void __launch_bounds__(Pow2Tile<L, C, STRIDED>::THREADS, Pow2Tile<L, C, STRIDED>::MIN_CTAS)
fft_pow2_ext_kernel(const FftPassArgs a) {
    extern __shared__ double2 sm[];
    for (int64_t blk = blockIdx.x; blk < a.ntiles; blk += gridDim.x) {
        // reverse: walk the tiles from the end (the data the previous pass wrote last is still in L2)
        fft_pow2_tile<L, C, STRIDED, true>(a, a.reverse ? a.ntiles - 1 - blk : blk, sm);
        if (blk + gridDim.x < a.ntiles) {
            if (!STRIDED && a.ls_bulk) {  // the issuing threads: the bulk copies have read their shared-memory source
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        }
    }
    if (!STRIDED && a.ls_bulk) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // writes performed before the CTA retires
    }
}

// The hot path: one tile per CTA, natural layout.
template <int L, int C, bool STRIDED>
__global__ void __launch_bounds__(Pow2Tile<L, C, STRIDED>::THREADS, Pow2Tile<L, C, STRIDED>::MIN_CTAS)
fft_pow2_kernel(const FftPassArgs a) {
    extern __shared__ double2 sm[];
    fft_pow2_tile<L, C, STRIDED, false>(a, blockIdx.x, sm);
}

// The same with the L2 prefetch of a tile further on (FftPassArgs::prefetch_ahead); a separate instantiation because the
// extra address arithmetic does not fit the hot path's 64 registers without spilling.
template <int L, int C>
__global__ void __launch_bounds__(Pow2Tile<L, C, true>::THREADS, Pow2Tile<L, C, true>::MIN_CTAS)
fft_pow2_pf_kernel(const FftPassArgs a) {
    extern __shared__ double2 sm[];
    fft_pow2_tile<L, C, true, false, true>(a, blockIdx.x, sm);
}

// ---------------------------------------------------------------------------------------------------------
// The exchange pass of the slab transform as a software-pipelined persistent kernel (line scatter only, see
// FftPassArgs::ls_*): one elected thread per CTA drives the SM's bulk-copy unit -- cp.async.bulk global -> shared for
// the NEXT tile's lines (completion on an mbarrier) and cp.async.bulk shared -> global for the PREVIOUS tile's
// transformed lines (into the peers' memory over NVLink) -- while all threads run the butterflies of the CURRENT
// tile. Three tile buffers rotate: a buffer is, in turn, the landing zone of the bulk load (dense lines), the
// exchange area of the Stockham stages (padded lines), and the source of the bulk stores (dense lines again).
// With loads and stores off the threads' critical path the pass is bound by the link, not by the latency of one
// tile's load -> compute -> store chain, which is what it degrades to (500 instead of 690 GB/s, measured on 8 B200)
// when the x pass of the previous group shares the SMs and HBM with it.
// ---------------------------------------------------------------------------------------------------------
template <int L, int C>
struct ExchangeTile {
    static constexpr int T = L / 8;
    static constexpr int THREADS = C * T;
    static constexpr int LINE = fft_pad(L);
    static constexpr int BUF_ELEMS = LINE * C;       // padded: the stages' layout is the larger one
    static constexpr int NBUF = 3;
    static constexpr int SMEM_BYTES = NBUF * BUF_ELEMS * 16 + 64;  // + the mbarriers
};

// source line and destination of line number `line` of this launch (the addressing of fft_pow2_tile's line scatter)
__device__ __forceinline__ void exchange_line(const FftPassArgs &a, int L, int64_t line, const double2 *&src, double2 *&dst) {
    if (a.ls_sub > 0) {
        const int64_t yl = line % a.ls_sub, t1 = line / a.ls_sub;
        const int q = (int) (t1 % a.ls_nblk);
        const int64_t xl = a.ls_first + t1 / a.ls_nblk;
        src = a.in + (((int64_t) q * a.ls_block_lines + xl) * a.ls_sub + yl) * L;
        dst = a.blk[q] + (xl * a.ls_plane_lines + yl) * L;
    } else {
        const int64_t yq = line % a.ls_group, t1 = line / a.ls_group;
        const int q = (int) (t1 % a.ls_nblk);
        const int64_t xl = t1 / a.ls_nblk;
        const int64_t yl = a.ls_first + yq;
        src = a.in + (xl * a.ls_plane_lines + (int64_t) q * a.ls_block_lines + yl) * L;
        dst = a.blk[q] + (xl * a.ls_block_lines + yl) * L;
    }
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    }
}

template <int L, int C>
__global__ void __launch_bounds__(ExchangeTile<L, C>::THREADS, 2) fft_exchange_kernel(const FftPassArgs a) {
    using Tile = ExchangeTile<L, C>;
    constexpr int T = Tile::T;
    extern __shared__ __align__(128) double2 sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + Tile::NBUF * Tile::BUF_ELEMS);
    const int tid = threadIdx.x, t = tid % T, c = tid / T;
    const int64_t ntiles = a.ntiles;

    // thread 0: arm the buffer's mbarrier and pull the C lines of tile `tile` into buffer `b`
    auto issue_load = [&](int64_t tile, int b) {
        const uint32_t bar = (uint32_t) __cvta_generic_to_shared(&bars[b]);
        const int64_t first = tile * C;
        const int nl = (int) (a.outer - first < C ? a.outer - first : C);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (nl * L * 16)) : "memory");
        for (int k = 0; k < nl; k++) {
            const double2 *src;
            double2 *dst;
            exchange_line(a, L, first + k, src, dst);
            const uint32_t sa = (uint32_t) __cvta_generic_to_shared(sm + b * Tile::BUF_ELEMS + k * L);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa),
                         "l"(src), "r"((uint32_t) (L * 16)), "r"(bar)
                         : "memory");
        }
    };

    if (tid == 0) {
        for (int b = 0; b < Tile::NBUF; b++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&bars[b])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int64_t) blockIdx.x < ntiles) {
            issue_load(blockIdx.x, 0);
        }
    }
    __syncthreads();

    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int b = (int) (it % Tile::NBUF);
        double2 *buf = sm + b * Tile::BUF_ELEMS;
        const int64_t line = tile * C + c;
        const bool active = line < a.outer;
        if (tid == 0 && tile + gridDim.x < ntiles) {
            // the buffer the next tile lands in was the store source two tiles ago: all but the latest store group
            // must have finished reading shared memory
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue_load(tile + gridDim.x, (int) ((it + 1) % Tile::NBUF));
        }
        mbar_wait((uint32_t) __cvta_generic_to_shared(&bars[b]), (uint32_t) ((it / Tile::NBUF) & 1));
        double2 v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const double2 z = active ? buf[c * L + t + m * T] : make_double2(0.0, 0.0);
            v[m] = a.inverse ? make_double2(z.y, z.x) : z;
        }
        __syncthreads();  // the dense lines have been read: the buffer becomes the stages' (padded) exchange area
        Pow2Stages<L, C, false, 1>::run(v, t, c, buf, a.tw);
        __syncthreads();  // the last exchange has been read: the buffer becomes the stores' (dense) source
        const double s = a.scale;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            buf[c * L + t + m * T] = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const int64_t first = tile * C;
            const int nl = (int) (a.outer - first < C ? a.outer - first : C);
            for (int k = 0; k < nl; k++) {
                const double2 *src;
                double2 *dst;
                exchange_line(a, L, first + k, src, dst);
                const uint32_t sa = (uint32_t) __cvta_generic_to_shared(buf + k * L);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sa),
                             "r"((uint32_t) (L * 16))
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every line has been written before the CTA retires
    }
}

// ---------------------------------------------------------------------------------------------------------
// A strided pass with TMA tile movement (experiment): the (L x C) tile -- C adjacent columns, row stride
// `inner` -- is described by a 3-D tensor map (2C doubles x L rows x planes) and moved by cp.async.bulk.tensor (SASS
// UTMALDG / UTMASTG) in boxes of 256 rows: one elected thread issues the copies, completion arrives on an mbarrier, and
// the tile lands in shared memory in exactly the layout the Stockham stages exchange through (row-major, C columns).
// NBUF = 1: one tile per CTA turn, load -> stages -> store, as many CTAs per SM as shared memory allows (the register
// path's structure with the LSU taken out of the tile movement). NBUF = 3: persistent CTAs, the next tile is pulled in
// while the current one is transformed and the previous one drains (the exchange kernel's structure).
// ---------------------------------------------------------------------------------------------------------
template <int L, int C, int NBUF>
struct TmaTile {
    static constexpr int T = L / 8;
    static constexpr int THREADS = C * T;
    static constexpr int BUF_ELEMS = Pow2Tile<L, C, true>::LINE * C;  // the stages' layout pads narrow tiles (C < 8)
    static constexpr int BOX_ROWS = L < 256 ? L : 256;
    static constexpr int SMEM_BYTES = NBUF * BUF_ELEMS * 16 + 128;
    // as many CTAs per SM as shared memory allows, with the registers to match
    static constexpr int MIN_CTAS = (227 * 1024) / SMEM_BYTES < 1 ? 1 : ((227 * 1024) / SMEM_BYTES) * THREADS > 1024
                    ? (1024 / THREADS < 1 ? 1 : 1024 / THREADS) : (227 * 1024) / SMEM_BYTES;
};

__device__ __forceinline__ void tma_load_3d(void *smem, const void *map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         (uint32_t) __cvta_generic_to_shared(smem)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tma_store_3d(const void *map, int x, int y, int z, const void *smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x), "r"(y), "r"(z),
                 "r"((uint32_t) __cvta_generic_to_shared(smem))
                 : "memory");
}

template <int L, int C, int NBUF>
__global__ void __launch_bounds__(TmaTile<L, C, NBUF>::THREADS, TmaTile<L, C, NBUF>::MIN_CTAS) fft_pow2_tma_kernel(const __grid_constant__ CUtensorMap in_map,
        const __grid_constant__ CUtensorMap out_map, const FftPassArgs a) {
    using Tile = TmaTile<L, C, NBUF>;
    constexpr int T = Tile::T;
    extern __shared__ __align__(128) double2 sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + NBUF * Tile::BUF_ELEMS);
    const int tid = threadIdx.x, c = tid % C, t = tid / C;
    const int64_t tiles_per_plane = (a.inner + C - 1) / C;
    const int64_t ntiles = a.ntiles;

    auto issue_load = [&](int64_t tile, int b) {
        const uint32_t bar = (uint32_t) __cvta_generic_to_shared(&bars[b]);
        const int o = (int) (tile / tiles_per_plane);
        const int x = (int) ((tile - (int64_t) o * tiles_per_plane) * C * 2);  // in doubles
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (L * C * 16)) : "memory");
#pragma unroll
        for (int y = 0; y < L; y += Tile::BOX_ROWS) {
            tma_load_3d(sm + b * Tile::BUF_ELEMS + y * C, &in_map, x, y, o, bar);
        }
    };

    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&bars[b])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (NBUF > 1 && (int64_t) blockIdx.x < ntiles) {
            issue_load(blockIdx.x, 0);
        }
    }
    __syncthreads();

    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int b = (int) (it % NBUF);
        double2 *buf = sm + b * Tile::BUF_ELEMS;
        if (tid == 0) {
            if (NBUF > 1) {
                if (tile + gridDim.x < ntiles) {
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF > 2 ? 1 : 0) : "memory");
                    issue_load(tile + gridDim.x, (int) ((it + 1) % NBUF));
                }
            } else {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's store has left the buffer
                issue_load(tile, 0);
            }
        }
        mbar_wait((uint32_t) __cvta_generic_to_shared(&bars[b]), (uint32_t) ((it / NBUF) & 1));
        double2 v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const double2 z = buf[(t + m * T) * C + c];
            v[m] = a.inverse ? make_double2(z.y, z.x) : z;
        }
        __syncthreads();  // the tile has been read: the buffer becomes the stages' exchange area (same layout)
        Pow2Stages<L, C, true, 1>::run(v, t, c, buf, a.tw);
        __syncthreads();
        const double s = a.scale;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            buf[(t + m * T) * C + c] = a.inverse ? make_double2(v[m].y * s, v[m].x * s) : make_double2(v[m].x * s, v[m].y * s);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const int o = (int) (tile / tiles_per_plane);
            const int x = (int) ((tile - (int64_t) o * tiles_per_plane) * C * 2);
#pragma unroll
            for (int y = 0; y < L; y += Tile::BOX_ROWS) {
                tma_store_3d(&out_map, x, y, o, buf + y * C);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// Real transforms of even power-of-two length N = 2H on the last axis, as ONE complex transform of length H
// (the packing trick: z[j] = x[2j] + i x[2j+1], which is exactly how a real line reads when viewed as double2):
//   r2c   Z = FFT_H(z);  X[k] = (Z[k] + conj Z[H-k])/2 - (i/2) w^k (Z[k] - conj Z[H-k]),  w = e^{-2 pi i/N},
//         k = 0..H-1, and X[H] = Re Z[0] - Im Z[0]; partners meet through one more shared-memory exchange.
//   c2r   Z'[k] = (X[k] + conj X[H-k]) + i w^{-k} (X[k] - conj X[H-k]);  (x[2j], x[2j+1]) = IFFT_H(Z')[j]
//         (the factor 2 of Z' = 2Z is the H-vs-N normalisation; `scale` carries 1/N as before).
// Half the butterflies, half the shared memory and no padding loads compared with transforming a full complex
// line per real line (what JavaFftService.rfft does on the CPU: JavaFftService.java:53-64). FUSE adds the fused
// convolution's product-on-load and crop-on-store (FftPassArgs::mul / crop_cols).
// ---------------------------------------------------------------------------------------------------------
template <int H, int C, bool FUSE>
__global__ void __launch_bounds__(Pow2Tile<H, C, false>::THREADS, Pow2Tile<H, C, false>::MIN_CTAS)
fft_real_pow2_kernel(const FftPassArgs a) {
    extern __shared__ double2 sm[];
    using Stages = Pow2Stages<H, C, false, 1>;
    constexpr int T = H / 8;
    const int tid = threadIdx.x, t = tid % T, c = tid / T;
    const int64_t line = (int64_t) blockIdx.x * C + c;
    const bool active = line < a.outer;
    const double2 *__restrict__ tw2 = a.tw2;  // tw2[k] = exp(-2 pi i k / (2H)), k in [0, H)
    const double s = a.scale;
    double2 v[8];
    if (a.mode == MODE_R2C) {
        const double2 *zin = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(a.in) + line * (2 * H));
#pragma unroll
        for (int m = 0; m < 8; m++) {
            v[m] = active ? zin[t + m * T] : make_double2(0.0, 0.0);
        }
        Stages::run(v, t, c, sm, a.tw);
        __syncthreads();  // the last exchange of the stages has been read by everyone
#pragma unroll
        for (int m = 0; m < 8; m++) {
            sm[Stages::sidx(t + m * T, c)] = v[m];
        }
        __syncthreads();
        if (!active) {
            return;
        }
        double2 *hout = a.out + line * (H + 1);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int k = t + m * T;
            const double2 zp = sm[Stages::sidx((H - k) & (H - 1), c)], w = __ldg(&tw2[k]);
            const double ex = 0.5 * (v[m].x + zp.x), ey = 0.5 * (v[m].y - zp.y);  // (Z + conj Zp) / 2
            const double ox = 0.5 * (v[m].y + zp.y), oy = -0.5 * (v[m].x - zp.x);  // (-i/2) (Z - conj Zp)
            hout[k] = make_double2((ex + (ox * w.x - oy * w.y)) * s, (ey + (ox * w.y + oy * w.x)) * s);
        }
        if (t == 0) {
            hout[H] = make_double2((v[0].x - v[0].y) * s, 0.0);
        }
    } else {  // MODE_C2R
        const int64_t hoff = (FUSE && a.crop_cols > 0 ? crop_source_line(a, line) : line) * (H + 1);
        if (active) {
            const double2 *hin = a.in + hoff;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int k = t + m * T;
                double2 x = hin[k], xp = hin[H - k];
                if (FUSE && a.mul) {
                    x = cmul(x, a.mul[hoff + k]);
                    xp = cmul(xp, a.mul[hoff + H - k]);
                }
                const double2 w = __ldg(&tw2[k]);
                const double sx = x.x + xp.x, sy = x.y - xp.y;  // X + conj Xp
                const double dx = x.x - xp.x, dy = x.y + xp.y;  // X - conj Xp
                // i * conj(w) * d = i * (w.x - i w.y) * (dx + i dy)
                const double px = w.x * dx + w.y * dy, py = w.x * dy - w.y * dx;  // conj(w) * d
                v[m] = make_double2(sy + px, sx - py);  // swap(Z'), Z' = (sx - py, sy + px): backward via the swap identity
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                v[m] = make_double2(0.0, 0.0);
            }
        }
        Stages::run(v, t, c, sm, a.tw);
        if (!active) {
            return;
        }
        if (FUSE && a.crop_cols > 0) {
            double *rout = reinterpret_cast<double *>(a.out) + line * a.crop_cols;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int j = 2 * (t + m * T);
                if (j < a.crop_cols) {
                    rout[j] = v[m].y * s;
                }
                if (j + 1 < a.crop_cols) {
                    rout[j + 1] = v[m].x * s;
                }
            }
        } else {
            double2 *zout = reinterpret_cast<double2 *>(reinterpret_cast<double *>(a.out) + line * (2 * H));
#pragma unroll
            for (int m = 0; m < 8; m++) {
                zout[t + m * T] = make_double2(v[m].y * s, v[m].x * s);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Any-length fallback kernel (still on the GPU; there is no CPU path): one line per CTA, Stockham autosort
// over the prime factors of L with O(p) work per output, ping-ponging between two shared-memory buffers.
// Exact DFT for every L that fits (2 * 16 * L bytes of shared memory); used for the non-power-of-two
// lengths the reference's own tests exercise (3, 5, 6, 7, 27: ArrayFftTest.java:61-224) and for L < 8.
// ---------------------------------------------------------------------------------------------------------
#define SSTC_FFT_MAX_FACTORS 32

struct FftGenericArgs {
    FftPassArgs p;
    int L;
    int nfactors;
    int factors[SSTC_FFT_MAX_FACTORS];
};

__global__ void __launch_bounds__(256) fft_generic_kernel(const FftGenericArgs g) {
    extern __shared__ double2 sm[];
    const FftPassArgs &a = g.p;
    const int L = g.L;
    double2 *cur = sm, *nxt = sm + L;

    // blockIdx -> line (o, i)
    const int64_t o = blockIdx.x / a.inner;
    const int64_t i = blockIdx.x - o * a.inner;
    const int64_t line = blockIdx.x;  // r2c / c2r are launched with inner == 1
    const int half = L / 2 + 1;
    const int64_t hoff = (a.crop_cols > 0 ? crop_source_line(a, line) : line) * half;  // c2r source line

    for (int p = threadIdx.x; p < L; p += blockDim.x) {
        double2 z;
        if (a.in_ksplit > 0) {
            const int q = p / a.in_ksplit;
            const int pk = p - q * a.in_ksplit;
            z = a.in_blk[q][(o * a.in_ksplit + pk) * a.in_blk_pitch + i];
        } else if (a.mode == MODE_C2C) {
            const int64_t idx = o * a.in_plane + (int64_t) p * a.in_kstride + i;
            z = a.in[idx];
            if (a.mul) {
                z = cmul(z, a.mul[idx]);
            }
        } else if (a.mode == MODE_R2C) {
            z = make_double2(reinterpret_cast<const double *>(a.in)[line * L + p], 0.0);
        } else {
            const int q = p < half ? p : L - p;
            z = a.in[hoff + q];
            if (a.mul) {
                z = cmul(z, a.mul[hoff + q]);
            }
            if (p >= half) {
                z.y = -z.y;
            }
        }
        cur[p] = a.inverse ? make_double2(z.y, z.x) : z;
    }
    __syncthreads();

    int ns = 1;
    for (int f = 0; f < g.nfactors; f++) {
        const int p = g.factors[f];
        const int lp = L / p;        // input stride between the p legs of a butterfly
        const int m = ns * p;
        const int tstep = L / m;     // twiddle exponent unit for w_{ns*p}
        for (int idx = threadIdx.x; idx < L; idx += blockDim.x) {
            const int k = idx % ns;
            const int q = (idx / ns) % p;
            const int jhi = idx / m;
            const int j = jhi * ns + k;
            // exponent advanced per leg r: r * (k*L/(ns*p) + q*L/p)  (mod L)
            const int step = (int) (((int64_t) k * tstep + (int64_t) q * lp) % L);
            int e = 0;
            double2 acc = cur[j];
            for (int r = 1; r < p; r++) {
                e += step;
                if (e >= L) {
                    e -= L;
                }
                const double2 w = __ldg(&a.tw[e]);
                const double2 x = cur[j + r * lp];
                acc.x += x.x * w.x - x.y * w.y;
                acc.y += x.x * w.y + x.y * w.x;
            }
            nxt[idx] = acc;
        }
        __syncthreads();
        double2 *tmp = cur;
        cur = nxt;
        nxt = tmp;
        ns = m;
    }

    const double s = a.scale;
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
        double2 z = cur[p];
        z = a.inverse ? make_double2(z.y * s, z.x * s) : make_double2(z.x * s, z.y * s);
        if (a.ksplit > 0) {
            const int q = p / a.ksplit;
            const int pk = p - q * a.ksplit;
            a.blk[q][(o * a.ksplit + pk) * a.blk_pitch + i] = z;
        } else if (a.mode == MODE_C2C) {
            a.out[o * a.out_plane + (int64_t) p * a.out_kstride + i] = z;
        } else if (a.mode == MODE_R2C) {
            if (p < half) {
                a.out[line * half + p] = z;
            }
        } else if (a.crop_cols > 0) {
            if (p < a.crop_cols) {
                reinterpret_cast<double *>(a.out)[line * a.crop_cols + p] = z.x;
            }
        } else {
            reinterpret_cast<double *>(a.out)[line * L + p] = z.x;
        }
    }
}

}  // namespace sstc

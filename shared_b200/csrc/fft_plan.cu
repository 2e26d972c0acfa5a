// fft_plan.cu -- FftService half of the C ABI: plans, twiddle tables, axis-pass dispatch, host tier.
//
// Replaces (paths relative to /root/reference):
//   src/org/shared/fft/JavaFftService.java:53-94 + FftOps.java:55-106   (pure-Java provider)
//   native/src/sharedx/Plan.cpp:44-141,230-437                            (FFTW-backed provider)
// Schedule for dims (d0..d_{r-1}), row-major: the last axis first (out of place, in -> out), then every
// other axis in place on `out`; the inverse applies 1/N in the last pass. The reference's per-dimension
// loop (FftOps.java:65-84) visits the same axes, rotating the layout each time instead.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include <algorithm>

#include "fft_aux_kernels.cuh"
#include "fft_internal.h"
#include "fft_kernels.cuh"
#include "sstc_internal.h"

namespace sstc {

// ---- twiddle tables -----------------------------------------------------------------------------------

static std::mutex g_tw_mutex;
static std::map<std::pair<int, int>, double2 *> g_tw_cache;  // (device, +-L) -> table (-L: stage-major pow2)

static double2 unit_root(int64_t num, int64_t den) {
    // exp(-2 pi i num/den), evaluated in long double then rounded (FftOps.java:155-168 uses double cos/sin
    // of a double phase; this is at least as accurate). Exact values stay exact.
    num %= den;
    if (num == 0) {
        return make_double2(1.0, 0.0);
    }
    if (4 * num == den) {
        return make_double2(0.0, -1.0);
    }
    if (2 * num == den) {
        return make_double2(-1.0, 0.0);
    }
    if (4 * num == 3 * den) {
        return make_double2(0.0, 1.0);
    }
    const long double two_pi = 6.283185307179586476925286766559005768L;
    long double ang = two_pi * (long double) num / (long double) den;
    return make_double2((double) cosl(ang), (double) -sinl(ang));
}

static const double2 *upload_table(int key_len, const std::vector<double2> &host) {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    double2 *d = nullptr;
    SSTC_CUDA(cudaMalloc(&d, sizeof(double2) * (host.empty() ? 1 : host.size())));
    if (!host.empty()) {
        SSTC_CUDA(cudaMemcpy(d, host.data(), sizeof(double2) * host.size(), cudaMemcpyHostToDevice));
    }
    g_tw_cache[std::make_pair(dev, key_len)] = d;
    return d;
}

// Full-circle table for the any-length kernel: tw[m] = exp(-2 pi i m / L).
static const double2 *twiddles_for(int L) {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto it = g_tw_cache.find(std::make_pair(dev, L));
    if (it != g_tw_cache.end()) {
        return it->second;
    }
    std::vector<double2> host((size_t) L);
    for (int m = 0; m < L; m++) {
        host[(size_t) m] = unit_root(m, L);
    }
    return upload_table(L, host);
}

// Stage-major table for the pow2 kernel (layout: pow2_twiddle_count in fft_kernels.cuh).
static const double2 *pow2_twiddles_for(int L) {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto it = g_tw_cache.find(std::make_pair(dev, -L));
    if (it != g_tw_cache.end()) {
        return it->second;
    }
    std::vector<double2> host;
    host.reserve((size_t) pow2_twiddle_count(L));
    for (int ns = pow2_stage_radix(L, 1); ns < L; ns *= pow2_stage_radix(L, ns)) {
        const int R = pow2_stage_radix(L, ns);
        for (int r = 1; r < R; r++) {
            for (int k = 0; k < ns; k++) {
                host.push_back(unit_root((int64_t) r * k, (int64_t) ns * R));
            }
        }
    }
    return upload_table(-L, host);
}

// Post-/pre-processing table of the packed real transforms: tw2[k] = exp(-2 pi i k / N), k in [0, N/2).
static const double2 *real_twiddles_for(int N) {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    const int key = N + (1 << 28);
    auto it = g_tw_cache.find(std::make_pair(dev, key));
    if (it != g_tw_cache.end()) {
        return it->second;
    }
    std::vector<double2> host((size_t) N / 2);
    for (int k = 0; k < N / 2; k++) {
        host[(size_t) k] = unit_root(k, N);
    }
    return upload_table(key, host);
}

// ---- axis-pass dispatch ------------------------------------------------------------------------------

constexpr int strided_cols(int L) { return L <= 64 ? 32 : L <= 256 ? 16 : L <= 512 ? 8 : L <= 1024 ? 4 : 2; }
constexpr int contig_lines(int L) { return (256 / (L / 8)) < 1 ? 1 : (256 / (L / 8)) > 64 ? 64 : (256 / (L / 8)); }

static const int64_t kMaxGrid = 2147483647LL;
static const int kGenericMaxL = 7168;  // two ping-pong buffers of 16*L bytes within 227 KB of shared memory

template <int L, int C, bool STRIDED>
static void launch_pow2(const FftPassArgs &a, int64_t grid, cudaStream_t stream) {
    using Tile = Pow2Tile<L, C, STRIDED>;
    static std::once_flag once[16];  // per device
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    if (Tile::SMEM_BYTES > 48 * 1024) {
        std::call_once(once[dev & 15], [] {
            SSTC_CUDA(cudaFuncSetAttribute(fft_pow2_kernel<L, C, STRIDED>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM_BYTES));
            SSTC_CUDA(cudaFuncSetAttribute(fft_pow2_ext_kernel<L, C, STRIDED>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM_BYTES));
        });
    }
    if (grid > kMaxGrid) {
        throw std::runtime_error("FFT batch too large for one launch");
    }
    // anything beyond "one tile per CTA, natural layout" runs the extended kernel
    const bool ext = a.ksplit > 0 || a.in_ksplit > 0 || a.ls_group > 0 || a.grid_limit > 0 || a.reverse || a.mul ||
            a.crop_cols > 0 || (STRIDED && (a.in_kstride != a.inner || a.out_kstride != a.inner));
    if (!ext) {
        if (STRIDED && L == 512 && a.prefetch_ahead > 0 && a.mode == MODE_C2C) {
            static std::once_flag pf_once[16];
            std::call_once(pf_once[dev & 15], [] {
                SSTC_CUDA(cudaFuncSetAttribute(fft_pow2_pf_kernel<L, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                        Pow2Tile<L, C, true>::SMEM_BYTES));
            });
            fft_pow2_pf_kernel<L, C><<<(unsigned) grid, Tile::THREADS, Tile::SMEM_BYTES, stream>>>(a);
            check_launch("fft_pow2_pf_kernel");
            return;
        }
        fft_pow2_kernel<L, C, STRIDED><<<(unsigned) grid, Tile::THREADS, Tile::SMEM_BYTES, stream>>>(a);
        check_launch("fft_pow2_kernel");
        return;
    }
    FftPassArgs b = a;
    b.ntiles = grid;
    if (a.grid_limit > 0 && a.grid_limit < grid) {
        grid = a.grid_limit;
    }
    fft_pow2_ext_kernel<L, C, STRIDED><<<(unsigned) grid, Tile::THREADS, Tile::SMEM_BYTES, stream>>>(b);
    check_launch("fft_pow2_ext_kernel");
}

// Diagnostics (sstc_exp_set "tma_strided"): 0 = off; 1 = the 512-point strided passes move their tiles with TMA, one
// tile per CTA turn, two CTAs per SM; 3 = persistent CTAs with three tile buffers (load / transform / store overlapped).
static std::atomic<int64_t> g_tma_strided{0};
// Diagnostics / tuning (sstc_exp_set "l2_prefetch"): prefetch distance in tiles for the strided hot path (0 = off).
static std::atomic<int64_t> g_l2_prefetch{0};
// Tuning (sstc_exp_set "long_strided_ctas"): strided passes whose tile fills an SM (L >= 2048: one 1024-thread CTA per SM)
// run as a persistent grid of this many CTAs looping over the tiles instead of one CTA per tile, which takes the CTA
// turnover (drain 1024 threads, schedule 1024 threads) out of every tile. 0 = one CTA per tile.
static std::atomic<int64_t> g_long_strided_ctas{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
        CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SSTC_CUDA(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) {
            throw std::runtime_error("cuTensorMapEncodeTiled is not available");
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// (outer, L, inner) complex128 viewed as a 3-D tensor of doubles: 2*inner x L x outer, boxes of 2C x box_rows x 1
static CUtensorMap strided_tile_map(const void *base, int64_t outer, int L, int64_t inner, int C, int box_rows) {
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t) inner * 2, (cuuint64_t) L, (cuuint64_t) outer};
    const cuuint64_t strides[2] = {(cuuint64_t) inner * 16, (cuuint64_t) L * (cuuint64_t) inner * 16};
    const cuuint32_t box[3] = {(cuuint32_t) (2 * C), (cuuint32_t) box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int) r) + ")");
    }
    return m;
}

template <int L, int C, int NBUF>
static void launch_tma(const FftPassArgs &a0, cudaStream_t stream) {
    using Tile = TmaTile<L, C, NBUF>;
    static std::once_flag once[16];
    static int sms[16];
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::call_once(once[dev & 15], [dev] {
        SSTC_CUDA(cudaFuncSetAttribute(fft_pow2_tma_kernel<L, C, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile::SMEM_BYTES));
        SSTC_CUDA(cudaDeviceGetAttribute(&sms[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    });
    FftPassArgs a = a0;
    const int64_t tiles_per_plane = (a.inner + C - 1) / C;
    a.ntiles = a.outer * tiles_per_plane;
    const CUtensorMap in_map = strided_tile_map(a.in, a.outer, L, a.inner, C, Tile::BOX_ROWS);
    const CUtensorMap out_map = strided_tile_map(a.out, a.outer, L, a.inner, C, Tile::BOX_ROWS);
    const int per_sm = Tile::MIN_CTAS;
    const int64_t grid = std::min<int64_t>(a.ntiles, (int64_t) sms[dev & 15] * per_sm);
    fft_pow2_tma_kernel<L, C, NBUF><<<(unsigned) grid, Tile::THREADS, Tile::SMEM_BYTES, stream>>>(in_map, out_map, a);
    check_launch("fft_pow2_tma_kernel");
}

template <int L>
static void launch_pow2_L(const FftPassArgs &a, bool strided, cudaStream_t stream) {
    if (L == 512) {
        // Tile-width experiments for the headline length (env SSTC_FFT512_{STRIDED,CONTIG}_C); the defaults
        // below are the measured winners.
        static const int sc = getenv("SSTC_FFT512_STRIDED_C") ? atoi(getenv("SSTC_FFT512_STRIDED_C")) : 0;
        static const int cc = getenv("SSTC_FFT512_CONTIG_C") ? atoi(getenv("SSTC_FFT512_CONTIG_C")) : 0;
        if (strided && (sc == 4 || sc == 16)) {
            int64_t tiles = (a.inner + sc - 1) / sc;
            if (sc == 4) {
                launch_pow2<512, 4, true>(a, a.outer * tiles, stream);
            } else {
                launch_pow2<512, 16, true>(a, a.outer * tiles, stream);
            }
            return;
        }
        if (!strided && (cc == 2 || cc == 8)) {
            if (cc == 2) {
                launch_pow2<512, 2, false>(a, (a.outer + 1) / 2, stream);
            } else {
                launch_pow2<512, 8, false>(a, (a.outer + 7) / 8, stream);
            }
            return;
        }
    }
    const int64_t tma = g_tma_strided.load();
    if (strided && tma > 0 && (L == 512 || L == 4096) && a.mode == MODE_C2C && a.ksplit == 0 && a.in_ksplit == 0 && !a.mul &&
            a.crop_cols == 0 && a.in_kstride == a.inner && a.out_kstride == a.inner && a.in_plane == (int64_t) L * a.inner &&
            a.out_plane == (int64_t) L * a.inner && a.inner * 2 < (1LL << 31) && a.outer < (1LL << 31)) {
        if (L == 512) {
            if (tma >= 3) {
                launch_tma<512, strided_cols(512), 3>(a, stream);
            } else if (tma == 2) {
                launch_tma<512, strided_cols(512), 2>(a, stream);
            } else {
                launch_tma<512, strided_cols(512), 1>(a, stream);
            }
        } else {
            launch_tma<4096, strided_cols(4096), 1>(a, stream);
        }
        return;
    }
    if (strided) {
        constexpr int C = strided_cols(L);
        int64_t tiles = (a.inner + C - 1) / C;
        launch_pow2<L, C, true>(a, a.outer * tiles, stream);
    } else {
        constexpr int C = contig_lines(L);
        launch_pow2<L, C, false>(a, (a.outer + C - 1) / C, stream);
    }
}

// Packed real transforms (r2c / c2r of even power-of-two length N = 2H, 32 <= N <= 8192).
template <int H, bool FUSE>
static void launch_real_H(const FftPassArgs &a, cudaStream_t stream) {
    constexpr int C = contig_lines(H);
    using Tile = Pow2Tile<H, C, false>;
    static std::once_flag once[16];
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    if (Tile::SMEM_BYTES > 48 * 1024) {
        std::call_once(once[dev & 15], [] {
            SSTC_CUDA(cudaFuncSetAttribute(fft_real_pow2_kernel<H, C, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                    Tile::SMEM_BYTES));
        });
    }
    const int64_t grid = (a.outer + C - 1) / C;
    if (grid > kMaxGrid) {
        throw std::runtime_error("FFT batch too large for one launch");
    }
    fft_real_pow2_kernel<H, C, FUSE><<<(unsigned) grid, Tile::THREADS, Tile::SMEM_BYTES, stream>>>(a);
    check_launch("fft_real_pow2_kernel");
}

template <int H>
static void launch_real(const FftPassArgs &a, cudaStream_t stream) {
    if (a.mul || a.crop_cols > 0) {
        launch_real_H<H, true>(a, stream);
    } else {
        launch_real_H<H, false>(a, stream);
    }
}

static bool real_kernel_holds(int N) { return (N & (N - 1)) == 0 && N >= 32 && N <= 8192; }

static void dispatch_real(FftPassArgs &a, int N, cudaStream_t stream) {
    a.tw = pow2_twiddles_for(N / 2);
    a.tw2 = real_twiddles_for(N);
    switch (N / 2) {
    case 16: launch_real<16>(a, stream); break;
    case 32: launch_real<32>(a, stream); break;
    case 64: launch_real<64>(a, stream); break;
    case 128: launch_real<128>(a, stream); break;
    case 256: launch_real<256>(a, stream); break;
    case 512: launch_real<512>(a, stream); break;
    case 1024: launch_real<1024>(a, stream); break;
    case 2048: launch_real<2048>(a, stream); break;
    case 4096: launch_real<4096>(a, stream); break;
    default: throw std::runtime_error("internal: real transform length");
    }
}

// Diagnostics (sstc_exp_set "exchange_smem_kb"): dynamic shared memory requested by the exchange kernels beyond what they
// use, to keep other kernels' CTAs off the SMs they run on (tools/exp_overlap.py).
static std::atomic<int64_t> g_exchange_smem_kb{0};

// The software-pipelined exchange pass (fft_exchange_kernel): persistent grid, at most `grid_limit` CTAs (default: two
// per SM, what its three tile buffers allow at L = 512).
template <int L>
static void launch_exchange(const FftPassArgs &a0, cudaStream_t stream) {
    constexpr int C = contig_lines(L);
    using Tile = ExchangeTile<L, C>;
    static std::once_flag once[16];
    static int sms[16];
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::call_once(once[dev & 15], [dev] {
        SSTC_CUDA(cudaFuncSetAttribute(fft_exchange_kernel<L, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SSTC_CUDA(cudaDeviceGetAttribute(&sms[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    });
    const size_t smem = std::max<size_t>(Tile::SMEM_BYTES, std::min<size_t>(227, (size_t) g_exchange_smem_kb.load()) * 1024);
    FftPassArgs a = a0;
    a.ntiles = (a.outer + C - 1) / C;
    int64_t grid = a.grid_limit > 0 ? a.grid_limit : 2 * (int64_t) sms[dev & 15];
    grid = std::max<int64_t>(1, std::min(grid, a.ntiles));
    fft_exchange_kernel<L, C><<<(unsigned) grid, Tile::THREADS, smem, stream>>>(a);
    check_launch("fft_exchange_kernel");
}

static void dispatch_exchange(const FftPassArgs &a, int L, cudaStream_t stream) {
    switch (L) {
    case 8: launch_exchange<8>(a, stream); break;
    case 16: launch_exchange<16>(a, stream); break;
    case 32: launch_exchange<32>(a, stream); break;
    case 64: launch_exchange<64>(a, stream); break;
    case 128: launch_exchange<128>(a, stream); break;
    case 256: launch_exchange<256>(a, stream); break;
    case 512: launch_exchange<512>(a, stream); break;
    case 1024: launch_exchange<1024>(a, stream); break;
    case 2048: launch_exchange<2048>(a, stream); break;
    case 4096: launch_exchange<4096>(a, stream); break;
    default: throw std::runtime_error("internal: exchange length");
    }
}

static std::vector<int> prime_factors(int n) {
    std::vector<int> f;
    for (int p = 2; (int64_t) p * p <= n; p++) {
        while (n % p == 0) {
            f.push_back(p);
            n /= p;
        }
    }
    if (n > 1) {
        f.push_back(n);
    }
    return f;
}

static void launch_generic(const FftPassArgs &a, int L, cudaStream_t stream) {
    if (L > kGenericMaxL) {
        throw std::runtime_error("FFT length " + std::to_string(L) + " is not supported by the CUDA provider");
    }
    FftGenericArgs g;
    g.p = a;
    g.L = L;
    std::vector<int> f = prime_factors(L);
    if ((int) f.size() > SSTC_FFT_MAX_FACTORS) {
        throw std::runtime_error("FFT length has too many factors");
    }
    g.nfactors = (int) f.size();
    for (size_t i = 0; i < f.size(); i++) {
        g.factors[i] = f[i];
    }
    size_t smem = (size_t) 2 * 16 * (size_t) L;
    static std::once_flag once[16];
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::call_once(once[dev & 15], [] {
        SSTC_CUDA(cudaFuncSetAttribute(fft_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                2 * 16 * kGenericMaxL));
    });
    int64_t lines = a.outer * a.inner;
    if (lines > kMaxGrid) {
        throw std::runtime_error("FFT batch too large for one launch");
    }
    int threads = L >= 256 ? 256 : L >= 64 ? 64 : 32;
    fft_generic_kernel<<<(unsigned) lines, threads, smem, stream>>>(g);
    check_launch("fft_generic_kernel");
}

static void dispatch_axis(const FftPassArgs &a, int L, bool strided, cudaStream_t stream);
static void long_axis(const void *in, void *out, int64_t outer, int L, int64_t inner, int mode, int inverse,
        double scale, cudaStream_t stream);

// Optional blocked addressing of one side of a pass (see FftPassArgs).
struct Blocks {
    int n = 0;
    void *const *ptrs = nullptr;
    int64_t pitch = 0;  // 0 = dense (== inner)
};

// Work folded into a pass by the fused convolution (see FftPassArgs::mul / crop_cols).
struct Fuse {
    const double2 *mul = nullptr;
    int crop_cols = 0, crop_nd = 0;
    int32_t crop_in[4] = {1, 1, 1, 1}, crop_out[4] = {1, 1, 1, 1};
};

bool tile_kernels_hold(int L) {
    const bool is_pow2 = (L & (L - 1)) == 0;
    return is_pow2 ? L <= 4096 : L <= kGenericMaxL;
}

// One pass over axis `L` of (outer, L, inner). For MODE_R2C / MODE_C2R, inner must be 1 and outer counts lines.
static void launch_axis(const void *in, void *out, int64_t outer, int L, int64_t inner, int mode, int inverse,
        double scale, cudaStream_t stream, const Blocks &bin = Blocks(), const Blocks &bout = Blocks(),
        const int64_t *layout = nullptr, int reverse = 0, const Fuse *fuse = nullptr) {
    if (outer <= 0 || inner <= 0 || L <= 0) {
        return;
    }
    // real transforms of power-of-two length: the packed half-length kernel (inputs / outputs are read as double2)
    const bool packed_real = mode != MODE_C2C && real_kernel_holds(L) && inner == 1 && bin.n == 0 && bout.n == 0 && !layout &&
            (mode == MODE_R2C ? !inverse : inverse) && !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) &&
            !(fuse && fuse->mul && (reinterpret_cast<uintptr_t>(fuse->mul) & 15));
    if (!packed_real && !tile_kernels_hold(L)) {
        // lengths the shared-memory kernels cannot hold: four-step (powers of two) or Bluestein (anything else)
        if (bin.n > 0 || bout.n > 0 || layout || fuse) {
            throw std::runtime_error("blocked / strided layouts need an axis length the tile kernels hold");
        }
        long_axis(in, out, outer, L, inner, mode, inverse, scale, stream);
        return;
    }
    FftPassArgs a;
    a.in = static_cast<const double2 *>(in);
    a.out = static_cast<double2 *>(out);
    const bool pow2 = L >= 8 && L <= 4096 && (L & (L - 1)) == 0;
    a.tw = packed_real ? nullptr : pow2 ? pow2_twiddles_for(L) : twiddles_for(L);
    a.tw2 = nullptr;
    a.outer = outer;
    a.inner = inner;
    a.in_plane = (int64_t) L * inner;
    a.out_plane = (int64_t) L * inner;
    a.in_kstride = inner;
    a.out_kstride = inner;
    if (layout) {
        a.in_plane = layout[0];
        a.in_kstride = layout[1];
        a.out_plane = layout[2];
        a.out_kstride = layout[3];
    }
    a.scale = scale;
    a.mode = mode;
    a.inverse = inverse;
    a.ksplit = a.in_ksplit = 0;
    a.ls_group = a.ls_nblk = a.ls_first = a.ls_block_lines = a.ls_plane_lines = a.ls_sub = a.ls_bulk = 0;
    a.ntiles = a.grid_limit = 0;
    a.reverse = reverse;
    a.prefetch_ahead = (int) g_l2_prefetch.load();
    if (L >= 2048 && inner > 1 && mode == MODE_C2C && g_long_strided_ctas.load() > 0) {
        a.grid_limit = g_long_strided_ctas.load();
    }
    a.mul = fuse ? fuse->mul : nullptr;
    a.crop_cols = fuse ? fuse->crop_cols : 0;
    a.crop_nd = fuse ? fuse->crop_nd : 0;
    for (int d = 0; d < 4; d++) {
        a.crop_in[d] = fuse ? fuse->crop_in[d] : 1;
        a.crop_out[d] = fuse ? fuse->crop_out[d] : 1;
    }
    a.blk_pitch = a.in_blk_pitch = inner;
    for (int q = 0; q < 8; q++) {
        a.blk[q] = nullptr;
        a.in_blk[q] = nullptr;
    }
    if (bin.n > 0 || bout.n > 0) {
        if (bin.n > 8 || bout.n > 8 || mode != MODE_C2C || (bin.n > 0 && (L % bin.n != 0 || !bin.ptrs)) ||
                (bout.n > 0 && (L % bout.n != 0 || !bout.ptrs)) || bin.pitch < 0 || bout.pitch < 0 ||
                (bin.pitch > 0 && bin.pitch < inner) || (bout.pitch > 0 && bout.pitch < inner)) {
            throw std::runtime_error("Invalid block specification");
        }
        if (bin.n > 0) {
            a.in_ksplit = L / bin.n;
            a.in_blk_pitch = bin.pitch > 0 ? bin.pitch : inner;
            for (int q = 0; q < bin.n; q++) {
                a.in_blk[q] = static_cast<const double2 *>(bin.ptrs[q]);
            }
        }
        if (bout.n > 0) {
            a.ksplit = L / bout.n;
            a.blk_pitch = bout.pitch > 0 ? bout.pitch : inner;
            for (int q = 0; q < bout.n; q++) {
                a.blk[q] = static_cast<double2 *>(bout.ptrs[q]);
            }
        }
    }
    if (packed_real) {
        dispatch_real(a, L, stream);
        return;
    }
    // blocked addressing lives in the STRIDED kernels; inner == 1 is routed there as a 1-column plane
    const bool strided = inner > 1 || bin.n > 0 || bout.n > 0 || layout;
    dispatch_axis(a, L, strided, stream);
}

static void dispatch_axis(const FftPassArgs &a, int L, bool strided, cudaStream_t stream) {
    switch (L) {
    case 8: launch_pow2_L<8>(a, strided, stream); break;
    case 16: launch_pow2_L<16>(a, strided, stream); break;
    case 32: launch_pow2_L<32>(a, strided, stream); break;
    case 64: launch_pow2_L<64>(a, strided, stream); break;
    case 128: launch_pow2_L<128>(a, strided, stream); break;
    case 256: launch_pow2_L<256>(a, strided, stream); break;
    case 512: launch_pow2_L<512>(a, strided, stream); break;
    case 1024: launch_pow2_L<1024>(a, strided, stream); break;
    case 2048: launch_pow2_L<2048>(a, strided, stream); break;
    case 4096: launch_pow2_L<4096>(a, strided, stream); break;
    default: launch_generic(a, L, stream); break;
    }
}

// ---- axes longer than the tile kernels hold ---------------------------------------------------------------
// Powers of two above 4096: four-step, L = L1*L2 (both <= 4096): sub-pass over n1 -> twiddle w_L^{n2 k1} +
// transpose -> sub-pass over n2; the result lands in natural order. Anything else above kGenericMaxL:
// Bluestein (chirp-z) over power-of-two transforms of size M >= 2L-1. Both need a workspace of the size of the
// data they touch, taken from per-thread grow-only scratch.

// 0: Bluestein lines, 1: four-step transpose, 2: real<->complex staging. Per host thread AND per stream: two streams
// driven by one thread may run their long-axis passes concurrently.
struct FftWorkspaces {
    Scratch ws[3];
};
static thread_local std::map<cudaStream_t, FftWorkspaces> t_fft_ws;
static const int kMaxAxisLog2 = 24;
static std::map<std::vector<int64_t>, double2 *> g_bluestein_filters;  // (device, L, M, inverse) -> FFT_M(b)

static int grid_for(int64_t items) {
    int64_t g = (items + 255) / 256;
    return (int) std::max<int64_t>(1, std::min<int64_t>(g, 148 * 16));
}

static void fourstep_axis(const double2 *in, double2 *out, int64_t outer, int L, int64_t inner, int inverse,
        double scale, cudaStream_t s) {
    int lg = 0;
    while ((1 << lg) < L) {
        lg++;
    }
    if (lg > kMaxAxisLog2) {
        throw std::runtime_error("FFT length " + std::to_string(L) + " exceeds the CUDA provider's 2^24 per-axis limit");
    }
    const int L1 = 1 << ((lg + 1) / 2), L2 = L / L1;
    const int64_t n = outer * (int64_t) L * inner;
    double2 *tmp = static_cast<double2 *>(t_fft_ws[s].ws[1].get(n * 16));
    launch_axis(in, tmp, outer, L1, (int64_t) L2 * inner, MODE_C2C, inverse, 1.0, s);
    if (inner == 1) {
        const int64_t tiles = outer * ((L1 + 31) / 32) * ((L2 + 31) / 32);
        fourstep_twiddle_transpose_tile_kernel<<<(int) std::min<int64_t>(tiles, 148 * 32), 256, 0, s>>>(tmp, out, outer, L1, L2,
                inverse);
    } else {
        fourstep_twiddle_transpose_kernel<<<grid_for(n), 256, 0, s>>>(tmp, out, outer, L1, L2, inner, inverse);
    }
    check_launch("fourstep_twiddle_transpose_kernel");
    launch_axis(out, out, outer, L2, (int64_t) L1 * inner, MODE_C2C, inverse, scale, s);
}

static const double2 *bluestein_filter(int64_t L, int64_t M, int inverse, cudaStream_t s) {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    const std::vector<int64_t> key{dev, L, M, inverse};
    {
        std::lock_guard<std::mutex> lock(g_tw_mutex);
        auto it = g_bluestein_filters.find(key);
        if (it != g_bluestein_filters.end()) {
            return it->second;
        }
    }
    double2 *B = nullptr;
    SSTC_CUDA(cudaMalloc(&B, (size_t) M * 16));
    bluestein_filter_kernel<<<grid_for(M), 256, 0, s>>>(B, L, M, inverse);
    check_launch("bluestein_filter_kernel");
    launch_axis(B, B, 1, (int) M, 1, MODE_C2C, 0, 1.0, s);
    // the table becomes visible to other threads / streams only once it is complete
    SSTC_CUDA(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto it = g_bluestein_filters.find(key);
    if (it != g_bluestein_filters.end()) {  // another thread built it meanwhile: keep theirs
        cudaFree(B);
        return it->second;
    }
    g_bluestein_filters[key] = B;
    return B;
}

static void bluestein_axis(const double2 *in, double2 *out, int64_t outer, int L, int64_t inner, int inverse,
        double scale, cudaStream_t s) {
    int64_t M = 1;
    while (M < 2 * (int64_t) L - 1) {
        M <<= 1;
    }
    if (M > (1LL << kMaxAxisLog2)) {
        throw std::runtime_error("FFT length " + std::to_string(L) + " exceeds the CUDA provider's per-axis limit");
    }
    const double2 *Bf = bluestein_filter(L, M, inverse, s);
    // whole planes (o) at a time, bounded to ~1 GiB of workspace
    const int64_t per_plane = inner * M * 16;
    const int64_t planes_per_chunk = std::max<int64_t>(1, (1LL << 30) / per_plane);
    for (int64_t o0 = 0; o0 < outer; o0 += planes_per_chunk) {
        const int64_t no = std::min(planes_per_chunk, outer - o0);
        const int64_t lines = no * inner;
        double2 *Y = static_cast<double2 *>(t_fft_ws[s].ws[0].get(lines * M * 16));
        const double2 *src = in + o0 * (int64_t) L * inner;
        double2 *dst = out + o0 * (int64_t) L * inner;
        bluestein_pre_kernel<<<grid_for(lines * M), 256, 0, s>>>(src, Y, no, L, inner, M, inverse);
        check_launch("bluestein_pre_kernel");
        launch_axis(Y, Y, lines, (int) M, 1, MODE_C2C, 0, 1.0, s);
        bluestein_mul_kernel<<<grid_for(lines * M), 256, 0, s>>>(Y, Bf, lines, M);
        check_launch("bluestein_mul_kernel");
        launch_axis(Y, Y, lines, (int) M, 1, MODE_C2C, 1, 1.0 / (double) M, s);
        bluestein_post_kernel<<<grid_for(no * (int64_t) L * inner), 256, 0, s>>>(Y, dst, no, L, inner, M, inverse, scale);
        check_launch("bluestein_post_kernel");
    }
}

static void long_axis(const void *in, void *out, int64_t outer, int L, int64_t inner, int mode, int inverse,
        double scale, cudaStream_t s) {
    const bool is_pow2 = (L & (L - 1)) == 0;
    auto c2c = [&](const double2 *src, double2 *dst, double sc) {
        if (is_pow2) {
            fourstep_axis(src, dst, outer, L, inner, inverse, sc, s);
        } else {
            bluestein_axis(src, dst, outer, L, inner, inverse, sc, s);
        }
    };
    if (mode == MODE_C2C) {
        c2c(static_cast<const double2 *>(in), static_cast<double2 *>(out), scale);
        return;
    }
    // real transforms on the last axis (inner == 1, outer = lines): stage through a full complex line
    const int64_t n = outer * (int64_t) L;
    double2 *full = static_cast<double2 *>(t_fft_ws[s].ws[2].get(n * 16));
    if (mode == MODE_R2C) {
        real_to_complex_kernel<<<grid_for(n), 256, 0, s>>>(static_cast<const double *>(in), full, n);
        check_launch("real_to_complex_kernel");
        c2c(full, full, scale);
        truncate_half_kernel<<<grid_for(n), 256, 0, s>>>(full, static_cast<double2 *>(out), outer, L);
        check_launch("truncate_half_kernel");
    } else {
        hermitian_expand_kernel<<<grid_for(n), 256, 0, s>>>(static_cast<const double2 *>(in), full, outer, L);
        check_launch("hermitian_expand_kernel");
        c2c(full, full, scale);
        complex_to_real_kernel<<<grid_for(n), 256, 0, s>>>(full, static_cast<double *>(out), n);
        check_launch("complex_to_real_kernel");
    }
}

// ---- plans ----------------------------------------------------------------------------------------------

}  // namespace sstc

namespace sstc {

void transform_lengths(int kind, int rank, const int32_t *dims, int64_t &in_len, int64_t &out_len) {
    // Plan::getTransformParameters, native/src/sharedx/Plan.cpp:230-320 (64-bit instead of jint).
    if (rank <= 0 || !dims) {
        throw std::runtime_error("Rank must be greater than zero");
    }
    for (int i = 0; i < rank; i++) {
        if (dims[i] <= 0) {
            throw std::runtime_error("Invalid dimensions");
        }
    }
    int64_t lead = 1;
    for (int i = 0; i < rank - 1; i++) {
        lead *= dims[i];
    }
    const int64_t last = dims[rank - 1];
    switch (kind) {
    case SSTC_FFT_R2C:
        in_len = lead * last;
        out_len = lead * 2 * (last / 2 + 1);
        break;
    case SSTC_FFT_C2R:
        in_len = lead * 2 * (last / 2 + 1);
        out_len = lead * last;
        break;
    case SSTC_FFT_FORWARD:
    case SSTC_FFT_BACKWARD:
        in_len = out_len = lead * 2 * last;
        break;
    default:
        throw std::runtime_error("Transform type not recognized");
    }
}

void build_plan(sstc_fft_plan_s &p) {
    p.steps.clear();
    if (p.workspace) {
        SSTC_CUDA(cudaFree(p.workspace));
        p.workspace = nullptr;
    }
    p.workspace_bytes = 0;
    const int rank = (int) p.dims.size();
    int64_t n = 1;
    for (int d : p.dims) {
        n *= d;
    }
    p.inverse = (p.kind == SSTC_FFT_BACKWARD || p.kind == SSTC_FFT_C2R) ? 1 : 0;
    const double inv_scale = p.inverse ? 1.0 / (double) n : 1.0;

    if (p.kind == SSTC_FFT_FORWARD || p.kind == SSTC_FFT_BACKWARD) {
        // last axis first (in -> out), then the others in place on out
        int64_t inner = 1;
        bool first = true;
        for (int a = rank - 1; a >= 0; a--) {
            const int L = p.dims[(size_t) a];
            if (L > 1 || (a == 0 && first)) {
                Step s{L, n / (L * inner), inner, MODE_C2C, first ? 0 : 1, 1, 1.0, {0, 0, 0, 0}};
                p.steps.push_back(s);
                first = false;
            }
            inner *= L;
        }
        p.steps.back().scale = inv_scale;
        // Transposed schedule for the two outermost axes. In place, the outermost pass loads AND stores
        // 128-byte row segments that are each in a different 2 MiB page (stride N/d0); measured at 512^3 that
        // costs 0.87 ms against 0.69 ms for the next axis. Instead the d1 pass stores its result transposed,
        // (d1, d0, rest), into a workspace, so the d0 pass loads with the small stride and stores back in
        // natural order: each pass has the page-hopping access on its (fire-and-forget) store side only.
        const size_t ns = p.steps.size();
        if (rank >= 3 && ns >= 3 && p.dims[0] > 1 && p.dims[1] > 1) {
            Step &s1 = p.steps[ns - 2], &s0 = p.steps[ns - 1];
            const int64_t d0 = p.dims[0], d1 = p.dims[1], rest = n / (d0 * d1);
            const bool big = rest * d1 * 16 >= (1 << 20) && n * 16 >= (256LL << 20);
            // measured at 512^3: the isolated passes gain (0.87 -> 0.75 ms) but the back-to-back step does not
            // (2.19 vs 2.18 ms), so automatic mode keeps the in-place schedule and its zero workspace.
            (void) big;
            const bool want = p.transposed_passes == 1;
            if (want && rest >= 8 && s1.axis_len == d1 && s0.axis_len == d0) {
                s1.dst = 2;
                s1.layout[0] = d1 * rest;  // read (x, y, i) natural
                s1.layout[1] = rest;
                s1.layout[2] = rest;       // write (y, x, i)
                s1.layout[3] = d0 * rest;
                s0.src = 2;
                s0.outer = d1;
                s0.inner = rest;
                s0.layout[0] = d0 * rest;  // read (y, x, i)
                s0.layout[1] = rest;
                s0.layout[2] = rest;       // write (x, y, i) natural
                s0.layout[3] = d1 * rest;
                p.workspace_bytes = n * 16;
            }
        }
    } else if (p.kind == SSTC_FFT_R2C) {
        const int last = p.dims[(size_t) rank - 1];
        const int64_t half = last / 2 + 1;
        p.steps.push_back(Step{last, n / last, 1, MODE_R2C, 0, 1, 1.0, {0, 0, 0, 0}});
        int64_t inner = half;
        for (int a = rank - 2; a >= 0; a--) {
            const int L = p.dims[(size_t) a];
            if (L > 1) {
                p.steps.push_back(
                        Step{L, (n / last) * half / (L * inner), inner, MODE_C2C, 1, 1, 1.0, {0, 0, 0, 0}});
            }
            inner *= L;
        }
    } else {  // C2R: leading axes on a workspace copy of the half-spectrum (the caller's `in` is preserved,
              // as the reference does for FFTW's input-destroying c2r: Plan.cpp:428-437), then c2r -> out
        const int last = p.dims[(size_t) rank - 1];
        const int64_t half = last / 2 + 1;
        const int64_t lines = n / last;
        int64_t inner = half;
        bool first = true;
        for (int a = rank - 2; a >= 0; a--) {
            const int L = p.dims[(size_t) a];
            if (L > 1) {
                p.steps.push_back(
                        Step{L, lines * half / (L * inner), inner, MODE_C2C, first ? 0 : 2, 2, 1.0, {0, 0, 0, 0}});
                first = false;
            }
            inner *= L;
        }
        if (!first) {
            p.workspace_bytes = lines * half * 16;
        }
        p.steps.push_back(Step{last, lines, 1, MODE_C2R, first ? 0 : 2, 1, inv_scale, {0, 0, 0, 0}});
    }
    if (p.workspace_bytes > 0) {
        SSTC_CUDA(cudaMalloc(&p.workspace, (size_t) p.workspace_bytes));
    }
    // Build every twiddle table now so execute() never allocates.
    for (const Step &s : p.steps) {
        const int L = s.axis_len;
        if (s.mode != MODE_C2C && real_kernel_holds(L)) {
            pow2_twiddles_for(L / 2);
            real_twiddles_for(L);
        }
        if (L >= 8 && L <= 4096 && (L & (L - 1)) == 0) {
            pow2_twiddles_for(L);
        } else if (L <= kGenericMaxL && (L & (L - 1)) != 0) {
            twiddles_for(L);
        } else if (L < 8) {
            twiddles_for(L);
        }  // longer axes build their tables (sub-pass twiddles, Bluestein filter) on first use
    }
}

void execute_plan(sstc_fft_plan_s &p, const double *d_in, double *d_out, cudaStream_t stream) {
    if (!d_in || !d_out) {
        throw std::runtime_error("Invalid arguments");
    }
    int pass = 0;
    for (const Step &s : p.steps) {
        const void *src = s.src == 0 ? (const void *) d_in : s.src == 1 ? (const void *) d_out : p.workspace;
        void *dst = s.dst == 1 ? (void *) d_out : p.workspace;
        // odd passes walk their tiles backwards: they begin on what the previous pass wrote last (L2-resident)
        launch_axis(src, dst, s.outer, s.axis_len, s.inner, s.mode, p.inverse, s.scale, stream, Blocks(), Blocks(),
                s.layout[1] ? s.layout : nullptr, (pass & 1) && p.alternate_order);
        pass++;
    }
}

static int plan_launches(const sstc_fft_plan_s &p) { return (int) p.steps.size(); }

static thread_local Scratch t_conv_ws[2];  // unfused fallback of the convolution: product, full-size result

// ConvolutionCache.convolve(ComplexArray cIm, RealArray ker) (ConvolutionCache.java:99-114) on a C2R plan:
// out = rifft(spec .* kernel_spec)[0:out_dims[0], 0:out_dims[1], ...]. Fused whenever every axis fits the tile
// kernels: the product is formed while the first pass loads, the crop happens while the last pass stores.
static void execute_convolve(sstc_fft_plan_s &p, const double *d_spec, const double *d_kernel_spec, double *d_out,
        const int32_t *out_dims, cudaStream_t stream) {
    if (p.kind != SSTC_FFT_C2R) {
        throw std::runtime_error("convolve needs a C2R plan");
    }
    if (!d_spec || !d_kernel_spec || !d_out || !out_dims) {
        throw std::runtime_error("Invalid arguments");
    }
    const int rank = (int) p.dims.size();
    int64_t out_lines = 1;
    for (int d = 0; d < rank; d++) {
        if (out_dims[d] <= 0 || out_dims[d] > p.dims[(size_t) d]) {
            throw std::runtime_error("Invalid kernel size");  // ConvolutionCache.java:80-81
        }
        if (d < rank - 1) {
            out_lines *= out_dims[d];
        }
    }
    bool fusable = rank - 1 <= 4;
    for (const Step &s : p.steps) {
        fusable = fusable && tile_kernels_hold(s.axis_len);
    }
    if (!fusable) {
        // product -> scratch, plain inverse -> scratch, strided copy of the valid block -> out
        const int64_t nspec = p.in_len, nfull = p.out_len;
        double *prod = static_cast<double *>(t_conv_ws[0].get(nspec * 8));
        double *full = static_cast<double *>(t_conv_ws[1].get(nfull * 8));
        if (sstc_eop_dev(2 /* CE_MUL */, SSTC_ELEM_COMPLEX, d_spec, d_kernel_spec, prod, nspec, stream) != 0) {
            throw std::runtime_error(std::string(sstc_last_error()));
        }
        execute_plan(p, prod, full, stream);
        std::vector<int32_t> bounds((size_t) 3 * rank), sD(p.dims), sS((size_t) rank), dD(out_dims, out_dims + rank),
                dS((size_t) rank);
        int64_t ss = 1, ds = 1, nout = 1;
        for (int d = rank - 1; d >= 0; d--) {
            sS[(size_t) d] = (int32_t) ss;
            dS[(size_t) d] = (int32_t) ds;
            ss *= p.dims[(size_t) d];
            ds *= out_dims[d];
            nout *= out_dims[d];
            bounds[(size_t) 3 * d] = 0;
            bounds[(size_t) 3 * d + 1] = 0;
            bounds[(size_t) 3 * d + 2] = out_dims[d];
        }
        if (sstc_map_dev(8, bounds.data(), 3 * rank, full, nfull, sD.data(), sS.data(), d_out, nout, dD.data(), dS.data(),
                    rank, stream) != 0) {
            throw std::runtime_error(std::string(sstc_last_error()));
        }
        return;
    }
    const size_t ns = p.steps.size();
    for (size_t k = 0; k < ns; k++) {
        const Step &s = p.steps[k];
        const void *src = s.src == 0 ? (const void *) d_spec : s.src == 1 ? (const void *) d_out : p.workspace;
        void *dst = s.dst == 1 ? (void *) d_out : p.workspace;
        Fuse f;
        bool any = false;
        int64_t outer = s.outer;
        if (k == 0) {
            f.mul = reinterpret_cast<const double2 *>(d_kernel_spec);
            any = true;
        }
        if (k == ns - 1) {  // the c2r pass
            f.crop_cols = out_dims[rank - 1];
            f.crop_nd = rank - 1;
            for (int d = 0; d < rank - 1; d++) {
                f.crop_in[d] = p.dims[(size_t) d];
                f.crop_out[d] = out_dims[d];
            }
            outer = out_lines;
            any = true;
        }
        launch_axis(src, dst, outer, s.axis_len, s.inner, s.mode, p.inverse, s.scale, stream, Blocks(), Blocks(), nullptr, 0,
                any ? &f : nullptr);
    }
}

sstc_fft_plan_s *new_plan(int kind, int rank, const int32_t *dims) {
    int64_t a = 0, b = 0;
    transform_lengths(kind, rank, dims, a, b);
    sstc_fft_plan_s *p = new sstc_fft_plan_s();
    p->kind = kind;
    p->dims.assign(dims, dims + rank);
    p->in_len = a;
    p->out_len = b;
    try {
        build_plan(*p);
    } catch (...) {
        if (p->workspace) {
            cudaFree(p->workspace);
        }
        delete p;
        throw;
    }
    return p;
}

void delete_plan(sstc_fft_plan_s *p) {
    if (p) {
        if (p->workspace) {
            cudaFree(p->workspace);
        }
        delete p;
    }
}

void fft_axis_pass(const void *in, void *out, int64_t outer, int L, int64_t inner, int mode, int inverse, double scale,
        cudaStream_t stream, const int64_t *layout) {
    static_assert(PASS_C2C == MODE_C2C && PASS_R2C == MODE_R2C && PASS_C2R == MODE_C2R, "pass modes");
    launch_axis(in, out, outer, L, inner, mode, inverse ? 1 : 0, scale, stream, Blocks(), Blocks(), layout);
}

static void scatter_common(FftPassArgs &a, const double *d_in, void *const *d_out_blocks, int nblocks, int32_t len,
        int inverse, double scale, int64_t max_ctas, int bulk) {
    if (!(len >= 8 && len <= 4096 && (len & (len - 1)) == 0)) {
        throw std::runtime_error("line scatter needs a power-of-two length between 8 and 4096");
    }
    memset(&a, 0, sizeof(a));
    a.in = reinterpret_cast<const double2 *>(d_in);
    a.tw = pow2_twiddles_for(len);
    a.inner = 1;
    a.in_plane = a.out_plane = len;
    a.in_kstride = a.out_kstride = 1;
    a.scale = scale;
    a.mode = MODE_C2C;
    a.inverse = inverse ? 1 : 0;
    a.blk_pitch = a.in_blk_pitch = 1;
    for (int q = 0; q < nblocks; q++) {
        a.blk[q] = static_cast<double2 *>(d_out_blocks[q]);
    }
    a.ls_nblk = nblocks;
    a.grid_limit = max_ctas > 0 ? max_ctas : 0;
    a.ls_bulk = bulk == 1 ? 1 : 0;
}

// bulk: 0 = per-thread stores, 1 = bulk stores from the tile kernel, 2 = the software-pipelined exchange kernel
static void dispatch_scatter(const FftPassArgs &a, int len, int bulk, cudaStream_t stream) {
    if (bulk == 2) {
        dispatch_exchange(a, len, stream);
    } else {
        dispatch_axis(a, len, false, stream);
    }
}

void fft_lines_scatter(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes, int32_t plane_lines,
        int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas, int bulk, cudaStream_t stream) {
    if (!d_in || !d_out_blocks || nblocks <= 0 || nblocks > 8 || planes < 0 || plane_lines <= 0 || len <= 0 ||
            plane_lines % nblocks != 0 || first < 0 || group <= 0 || first + group > plane_lines / nblocks) {
        throw std::runtime_error("Invalid arguments");
    }
    FftPassArgs a;
    scatter_common(a, d_in, d_out_blocks, nblocks, len, inverse, scale, max_ctas, bulk);
    if (planes == 0) {
        return;
    }
    a.outer = planes * nblocks * (int64_t) group;  // lines in this launch
    a.ls_group = group;
    a.ls_first = first;
    a.ls_block_lines = plane_lines / nblocks;
    a.ls_plane_lines = plane_lines;
    dispatch_scatter(a, len, bulk, stream);
}

void fft_planes_scatter(const double *d_in, void *const *d_out_blocks, int nblocks, int32_t block_planes, int32_t plane_lines,
        int32_t dst_plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas,
        int bulk, cudaStream_t stream) {
    if (!d_in || !d_out_blocks || nblocks <= 0 || nblocks > 8 || block_planes <= 0 || plane_lines <= 0 ||
            dst_plane_lines < plane_lines || len <= 0 || first < 0 || group <= 0 || first + group > block_planes) {
        throw std::runtime_error("Invalid arguments");
    }
    FftPassArgs a;
    scatter_common(a, d_in, d_out_blocks, nblocks, len, inverse, scale, max_ctas, bulk);
    a.outer = (int64_t) nblocks * group * plane_lines;  // lines in this launch
    a.ls_group = group;
    a.ls_first = first;
    a.ls_block_lines = block_planes;
    a.ls_plane_lines = dst_plane_lines;
    a.ls_sub = plane_lines;
    dispatch_scatter(a, len, bulk, stream);
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_fft_lengths(int kind, int rank, const int32_t *dims, int64_t *in_len, int64_t *out_len) {
    return guarded([&] {
        int64_t a = 0, b = 0;
        transform_lengths(kind, rank, dims, a, b);
        if (in_len) {
            *in_len = a;
        }
        if (out_len) {
            *out_len = b;
        }
    });
}

int sstc_fft_plan_create(sstc_fft_plan *plan, int kind, int rank, const int32_t *dims) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        *plan = nullptr;
        *plan = new_plan(kind, rank, dims);
    });
}

int sstc_fft_plan_execute(sstc_fft_plan plan, const double *d_in, double *d_out, void *stream) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        execute_plan(*plan, d_in, d_out, as_stream(stream));
    });
}

int sstc_fft_plan_destroy(sstc_fft_plan plan) {
    return guarded([&] {
        if (plan) {
            if (plan->workspace) {
                SSTC_CUDA(cudaFree(plan->workspace));
            }
            delete plan;
        }
    });
}

int64_t sstc_fft_plan_workspace(sstc_fft_plan plan) { return plan ? plan->workspace_bytes : 0; }

int sstc_fft_plan_launches(sstc_fft_plan plan) { return plan ? plan_launches(*plan) : 0; }

int sstc_fft_plan_set_hint(sstc_fft_plan plan, const char *name, int64_t value) {
    return guarded([&] {
        if (!plan || !name) {
            throw std::runtime_error("Invalid arguments");
        }
        if (std::string(name) == "alternate_order") {
            plan->alternate_order = value ? 1 : 0;
        } else if (std::string(name) == "transposed_passes") {
            if (value < -1 || value > 1) {
                throw std::runtime_error("Invalid hint value");
            }
            plan->transposed_passes = (int) value;
            build_plan(*plan);
        } else {
            throw std::runtime_error("Unknown hint");
        }
    });
}

int sstc_fft_convolve_dev(sstc_fft_plan plan, const double *d_spec, const double *d_kernel_spec, double *d_out,
        const int32_t *out_dims, void *stream) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        execute_convolve(*plan, d_spec, d_kernel_spec, d_out, out_dims, as_stream(stream));
    });
}

int sstc_fft_axis_dev(const double *d_in, double *d_out, int64_t outer, int32_t len, int64_t inner, int inverse,
        double scale, void *stream) {
    return guarded([&] {
        if (!d_in || !d_out || outer < 0 || inner < 0 || len <= 0) {
            throw std::runtime_error("Invalid arguments");
        }
        launch_axis(d_in, d_out, outer, len, inner, MODE_C2C, inverse ? 1 : 0, scale, as_stream(stream));
    });
}

int sstc_fft_axis_strided_dev(const double *d_in, double *d_out, int64_t outer, int32_t len, int64_t inner,
        int64_t in_plane, int64_t in_kstride, int64_t out_plane, int64_t out_kstride, int inverse, double scale,
        void *stream) {
    return guarded([&] {
        if (!d_in || !d_out || outer < 0 || inner < 0 || len <= 0) {
            throw std::runtime_error("Invalid arguments");
        }
        const int64_t layout[4] = {in_plane, in_kstride, out_plane, out_kstride};
        launch_axis(d_in, d_out, outer, len, inner, MODE_C2C, inverse ? 1 : 0, scale, as_stream(stream), Blocks(),
                Blocks(), layout);
    });
}

int sstc_fft_lines_scatter_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes,
        int32_t plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas,
        void *stream) {
    return guarded([&] {
        fft_lines_scatter(d_in, d_out_blocks, nblocks, planes, plane_lines, len, first, group, inverse, scale, max_ctas, 0,
                as_stream(stream));
    });
}

int sstc_exp_set(const char *name, int64_t value) {
    return guarded([&] {
        if (name && std::string(name) == "exchange_smem_kb") {
            g_exchange_smem_kb.store(value);
        } else if (name && std::string(name) == "tma_strided") {
            g_tma_strided.store(value);
        } else if (name && std::string(name) == "l2_prefetch") {
            g_l2_prefetch.store(value);
        } else if (name && std::string(name) == "long_strided_ctas") {
            g_long_strided_ctas.store(value);
        } else {
            throw std::runtime_error("Unknown option");
        }
    });
}

int sstc_fft_lines_scatter_ex_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes,
        int32_t plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas, int bulk,
        void *stream) {
    return guarded([&] {
        fft_lines_scatter(d_in, d_out_blocks, nblocks, planes, plane_lines, len, first, group, inverse, scale, max_ctas, bulk,
                as_stream(stream));
    });
}

int sstc_fft_planes_scatter_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int32_t block_planes,
        int32_t plane_lines, int32_t dst_plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale,
        int64_t max_ctas, void *stream) {
    return guarded([&] {
        fft_planes_scatter(d_in, d_out_blocks, nblocks, block_planes, plane_lines, dst_plane_lines, len, first, group,
                inverse, scale, max_ctas, 0, as_stream(stream));
    });
}

int sstc_fft_axis_blocked_dev(void *const *d_in_blocks, int n_in, int64_t in_pitch, void *const *d_out_blocks,
        int n_out, int64_t out_pitch, int64_t outer, int32_t len, int64_t inner, int inverse, double scale,
        void *stream) {
    return guarded([&] {
        if (!d_in_blocks || !d_out_blocks || n_in <= 0 || n_out <= 0 || outer < 0 || inner < 0 || len <= 0) {
            throw std::runtime_error("Invalid arguments");
        }
        Blocks bin, bout;
        bin.n = n_in;
        bin.ptrs = d_in_blocks;
        bin.pitch = in_pitch;
        bout.n = n_out;
        bout.ptrs = d_out_blocks;
        bout.pitch = out_pitch;
        launch_axis(nullptr, nullptr, outer, len, inner, MODE_C2C, inverse ? 1 : 0, scale, as_stream(stream), bin,
                bout);
    });
}

}  // extern "C"

// scan2d_core.cuh -- the inter-tile protocol of the single-pass 2-D sum scan (array_scan2d.cu; rdOp over both
// dimensions, DimensionOps.java:411-435, and createIntegralImage, NativeImageKernel.cpp:32-118): the order in which tiles
// are handed out, the descriptor words and the decoupled look-back. Written against a memory policy (relaxed 8-byte
// loads / stores of the descriptor words) so that the same code runs in the kernel (PTX ld/st.relaxed.gpu) and in
// tests/emu/scan2d_emu.cpp (host threads as CTAs, __atomic builtins, ThreadSanitizer): that the hand-out order cannot
// deadlock with fewer workers than tiles, and that a reader never takes a half-published descriptor, is checked there
// without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SSTC_HD __host__ __device__
#else
#define SSTC_HD
#endif

namespace sstc {

constexpr int kT = 64;  // tile edge

// One descriptor element: .x = the tile's own sum (`agg`: per row over its columns / per column of Y over its rows),
// .y = its inclusive prefix (`pre` = agg + everything left of / above the tile), side by side so that one 16-byte load
// reads both; each half is its own atomic word. The value IS the flag: the words start out as all-ones (a NaN no
// arithmetic produces: published NaNs are canonicalised), and an aligned 8-byte store is atomic, so a reader needs one
// round trip per predecessor and no fences.
struct alignas(16) ScanPair {
    double x, y;
};

struct ScanDesc {
    ScanPair ap[kT];
};

constexpr unsigned long long kNotReady = ~0ULL;
constexpr unsigned long long kCanonicalNaN = 0x7FF8000000000000ULL;

SSTC_HD inline unsigned long long scan_bits(double v) {
    union {
        double d;
        unsigned long long u;
    } c;
    c.d = v;
    return c.u;
}

SSTC_HD inline double scan_value(unsigned long long u) {
    union {
        double d;
        unsigned long long u;
    } c;
    c.u = u;
    return c.d;
}

template <class Mem>
SSTC_HD inline void publish(double *p, double v) {
    Mem::store(p, v != v ? kCanonicalNaN : scan_bits(v));  // never the "not ready" pattern
}

// Ticket -> tile position along anti-diagonals (d = tr + tc, tr ascending within a diagonal), in closed form: the
// diagonals grow by one tile up to min(R, C), stay at min(R, C) tiles up to max(R, C), then shrink. Both predecessors of
// a tile (left, above) lie on the previous diagonal, so they carry smaller tickets: every tile a worker may wait for
// was taken earlier by a worker that is running.
SSTC_HD inline void ticket_to_tile(unsigned int t, int R, int C, int &tr, int &tc) {
    const long long m = R < C ? R : C, M = R < C ? C : R, total = (long long) R * C;
    const long long head = m * (m + 1) / 2, body = (M - m) * m;
    long long d, first;
    if (t < head) {
        d = (long long) ((sqrt(8.0 * (double) t + 1.0) - 1.0) * 0.5);
        while (d * (d + 1) / 2 > t) {
            d--;
        }
        while ((d + 1) * (d + 2) / 2 <= t) {
            d++;
        }
        first = d * (d + 1) / 2;
    } else if (t < head + body) {
        d = m + (t - head) / m;
        first = head + (d - m) * m;
    } else {
        const long long u = total - 1 - t;  // mirrored: counts from the last tile
        long long e = (long long) ((sqrt(8.0 * (double) u + 1.0) - 1.0) * 0.5);
        while (e * (e + 1) / 2 > u) {
            e--;
        }
        while ((e + 1) * (e + 2) / 2 <= u) {
            e++;
        }
        d = (long long) R + C - 2 - e;
        first = total - (e + 1) * (e + 2) / 2;
    }
    const long long first_tr = d - (C - 1) > 0 ? d - (C - 1) : 0;
    tr = (int) (first_tr + (t - first));
    tc = (int) (d - tr);
}

// Sum of the predecessors' contributions for element `e` of the carry vector. `pos` walks towards the origin with
// stride `step` (1 for the tiles on the left, tiles_c for the tiles above); `count` predecessors exist. With the tiles
// handed out along anti-diagonals only the few diagonals in flight lack a prefix, so the walk is short. (A variant
// in which all 256 threads read four predecessors per round was measured and is SLOWER -- 0.35 against 0.30 ms at
// 8192 x 8192: the walk is already short, the extra barriers are not free.)
template <class Mem>
SSTC_HD inline double look_back(const ScanDesc *desc, int64_t pos, int64_t step, int count, int e) {
    double carry = 0.0;
    int64_t p = pos - step;
    for (int k = 0; k < count; k++, p -= step) {
        for (;;) {
            unsigned long long agg, pre;
            Mem::load_pair(&desc[p].ap[e], agg, pre);  // ONE round trip per predecessor
            if (pre != kNotReady) {
                return carry + scan_value(pre);
            }
            if (agg != kNotReady) {
                carry += scan_value(agg);
                break;
            }
        }
    }
    return carry;
}

}  // namespace sstc

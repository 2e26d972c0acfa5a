// TEST INFRASTRUCTURE ONLY -- the inter-tile protocol of the single-pass 2-D sum scan (shared_b200/csrc/scan2d_core.cuh:
// ticket order, descriptor words, decoupled look-back -- the code array_scan2d.cu's kernel runs) with host threads
// standing in for CTAs. A worker handles a tile serially (the intra-tile shuffles of the kernel are not what is checked
// here); what is checked is that any number of workers, down to ONE, finishes (a tile never waits for a ticket that has
// not been handed out), that no reader takes a half-published descriptor (ThreadSanitizer build), and that the result
// equals the serial summed-area table. Usage:
//   scan2d_emu scan <rows> <cols> <workers> <seed> <border 0|1> <inplace 0|1>
//   scan2d_emu tickets <max_r> <max_c>       every (R, C) up to the bounds: bijection, predecessors carry smaller tickets
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../shared_b200/csrc/scan2d_core.cuh"

using namespace sstc;

struct HostScanMem {
    static void load_pair(const ScanPair *p, unsigned long long &agg, unsigned long long &pre) {
        // the kernel reads both halves with one 16-byte load; two loads in either order see a state the walk handles
        agg = __atomic_load_n(reinterpret_cast<const unsigned long long *>(&p->x), __ATOMIC_RELAXED);
        pre = __atomic_load_n(reinterpret_cast<const unsigned long long *>(&p->y), __ATOMIC_RELAXED);
    }
    static void store(double *p, unsigned long long bits) {
        __atomic_store_n(reinterpret_cast<unsigned long long *>(p), bits, __ATOMIC_RELAXED);
    }
};

struct Job {
    const double *src;
    double *dst;
    int64_t rows, cols, spitch, dpitch;
    int tiles_r, tiles_c;
    bool border;
    ScanDesc *row_desc, *col_desc;
    unsigned int ticket;
};

static double input_at(const Job &j, int64_t gi, int64_t gj) {
    if (gi >= j.rows || gj >= j.cols) {
        return 0.0;
    }
    if (!j.border) {
        return j.src[gi * j.spitch + gj];
    }
    return (gi == 0 || gj == 0) ? j.dst[gi * j.dpitch + gj] : j.src[(gi - 1) * j.spitch + (gj - 1)];
}

static void *worker(void *arg) {
    Job &j = *static_cast<Job *>(arg);
    const int64_t ntiles = (int64_t) j.tiles_r * j.tiles_c;
    std::vector<double> tile((size_t) kT * kT);
    double row_carry[kT], own[kT];
    for (;;) {
        const unsigned int t = __atomic_fetch_add(&j.ticket, 1u, __ATOMIC_RELAXED);
        if ((int64_t) t >= ntiles) {
            return nullptr;
        }
        int tr, tc;
        ticket_to_tile(t, j.tiles_r, j.tiles_c, tr, tc);
        const int64_t pos = (int64_t) tr * j.tiles_c + tc, r0 = (int64_t) tr * kT, c0 = (int64_t) tc * kT;
        for (int r = 0; r < kT; r++) {  // load + row scan, the row's sum published at once
            double run = 0.0;
            for (int c = 0; c < kT; c++) {
                run += input_at(j, r0 + r, c0 + c);
                tile[(size_t) r * kT + c] = run;
            }
            publish<HostScanMem>(&j.row_desc[pos].ap[r].x, run);
            row_carry[r] = run;
        }
        for (int r = 0; r < kT; r++) {  // row carry
            const double c = look_back<HostScanMem>(j.row_desc, pos, 1, tc, r);
            publish<HostScanMem>(&j.row_desc[pos].ap[r].y, c + row_carry[r]);
            row_carry[r] = c;
        }
        for (int c = 0; c < kT; c++) {
            own[c] = 0.0;
        }
        for (int r = 0; r < kT; r++) {  // Y and its column sums
            for (int c = 0; c < kT; c++) {
                tile[(size_t) r * kT + c] += row_carry[r];
                own[c] += tile[(size_t) r * kT + c];
            }
        }
        for (int c = 0; c < kT; c++) {  // column carry, column scan, store
            publish<HostScanMem>(&j.col_desc[pos].ap[c].x, own[c]);
            const double cc = look_back<HostScanMem>(j.col_desc, pos, j.tiles_c, tr, c);
            publish<HostScanMem>(&j.col_desc[pos].ap[c].y, cc + own[c]);
            double run = cc;
            for (int r = 0; r < kT; r++) {
                run += tile[(size_t) r * kT + c];
                if (r0 + r < j.rows && c0 + c < j.cols) {
                    j.dst[(r0 + r) * j.dpitch + (c0 + c)] = run;
                }
            }
        }
    }
}

static int run_scan(int64_t rows, int64_t cols, int workers, unsigned seed, bool border, bool inplace) {
    // integer-valued data: every partial sum is exact, so any summation order gives the same bits
    const int64_t srows = border ? rows - 1 : rows, scols = border ? cols - 1 : cols;
    std::vector<double> src((size_t) (srows * scols)), dst((size_t) (rows * cols)), want((size_t) (rows * cols));
    auto next = [&]() {
        seed = seed * 1664525u + 1013904223u;
        return (double) ((int) ((seed >> 8) % 2001) - 1000);
    };
    for (auto &v : src) {
        v = next();
    }
    for (auto &v : dst) {
        v = next();  // border mode reads row 0 / column 0 of it; everything else must be overwritten
    }
    if (inplace && !border) {
        dst = src;
    }
    Job j;
    j.src = (inplace && !border) ? dst.data() : src.data();
    j.dst = dst.data();
    j.rows = rows;
    j.cols = cols;
    j.spitch = scols;
    j.dpitch = cols;
    j.border = border;
    j.tiles_r = (int) ((rows + kT - 1) / kT);
    j.tiles_c = (int) ((cols + kT - 1) / kT);
    j.ticket = 0;
    for (int64_t i = 0; i < rows; i++) {  // serial summed-area table of the virtual input
        double run = 0.0;
        for (int64_t c = 0; c < cols; c++) {
            run += input_at(j, i, c);
            want[(size_t) (i * cols + c)] = run + (i > 0 ? want[(size_t) ((i - 1) * cols + c)] : 0.0);
        }
    }
    const size_t nt = (size_t) j.tiles_r * j.tiles_c;
    std::vector<ScanDesc> desc(2 * nt);
    memset(desc.data(), 0xFF, desc.size() * sizeof(ScanDesc));
    j.row_desc = desc.data();
    j.col_desc = desc.data() + nt;
    std::vector<pthread_t> th((size_t) workers);
    for (auto &h : th) {
        pthread_create(&h, nullptr, worker, &j);
    }
    for (auto &h : th) {
        pthread_join(h, nullptr);
    }
    for (size_t i = 0; i < want.size(); i++) {
        if (memcmp(&dst[i], &want[i], 8) != 0) {
            fprintf(stderr, "mismatch at (%lld, %lld): %.17g != %.17g\n", (long long) (i / cols), (long long) (i % cols), dst[i],
                    want[i]);
            return 1;
        }
    }
    return 0;
}

static int run_tickets(int max_r, int max_c) {
    for (int R = 1; R <= max_r; R++) {
        for (int C = 1; C <= max_c; C++) {
            std::vector<int64_t> ticket_of((size_t) R * C, -1);
            for (unsigned int t = 0; t < (unsigned int) (R * C); t++) {
                int tr, tc;
                ticket_to_tile(t, R, C, tr, tc);
                if (tr < 0 || tr >= R || tc < 0 || tc >= C || ticket_of[(size_t) tr * C + tc] >= 0) {
                    fprintf(stderr, "R=%d C=%d ticket %u -> (%d, %d): out of range or taken twice\n", R, C, t, tr, tc);
                    return 1;
                }
                ticket_of[(size_t) tr * C + tc] = t;
            }
            for (int tr = 0; tr < R; tr++) {
                for (int tc = 0; tc < C; tc++) {
                    const int64_t t = ticket_of[(size_t) tr * C + tc];
                    if ((tr > 0 && ticket_of[(size_t) (tr - 1) * C + tc] >= t) || (tc > 0 && ticket_of[(size_t) tr * C + tc - 1] >= t)) {
                        fprintf(stderr, "R=%d C=%d tile (%d, %d): a predecessor carries a later ticket\n", R, C, tr, tc);
                        return 1;
                    }
                }
            }
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 4 && !strcmp(argv[1], "tickets")) {
        return run_tickets(atoi(argv[2]), atoi(argv[3]));
    }
    if (argc == 8 && !strcmp(argv[1], "scan")) {
        return run_scan(atoll(argv[2]), atoll(argv[3]), atoi(argv[4]), (unsigned) atoi(argv[5]), atoi(argv[6]) != 0, atoi(argv[7]) != 0);
    }
    return 2;
}

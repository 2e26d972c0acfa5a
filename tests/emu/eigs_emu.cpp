// TEST INFRASTRUCTURE ONLY -- drives shared_b200/csrc/eigs_core.cuh (the very code array_eigs.cu launches) on host
// threads. Usage: eigs_emu <n> <nthreads> <in.f64> <out.f64> [log_cap chunk_cap]; in = n*n doubles, out = n*n (vec) +
// 2n (val) doubles. Small log_cap / chunk_cap values exercise the log-full path and the chunked staging.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../shared_b200/csrc/eigs_core.cuh"
#include "cta_emu.h"

int main(int argc, char **argv) {
    if (argc != 5 && argc != 7) {
        return 2;
    }
    const int n = atoi(argv[1]), nthreads = atoi(argv[2]);
    std::vector<double> src((size_t) n * n), vec((size_t) n * n), val((size_t) 2 * n);
    FILE *f = fopen(argv[3], "rb");
    if (!f || fread(src.data(), 8, src.size(), f) != src.size()) {
        return 3;
    }
    fclose(f);
    // the device wrapper's layout: H padded to an odd row stride when it lives in shared memory
    const int ld = n | 1;
    std::vector<double> H((size_t) n * ld), V((size_t) n * n), ort((size_t) n + 1), T((size_t) n * n), X((size_t) n * n, 0.0);
    sstc::EigsShared sh;
    memset(&sh, 0, sizeof(sh));
    const int log_cap = argc == 7 ? atoi(argv[5]) : 8 * n * n + 16, chunk_cap = argc == 7 ? atoi(argv[6]) : 256;
    std::vector<sstc::EigsRot> log((size_t) log_cap), chunk((size_t) chunk_cap);
    std::vector<double> Vs((size_t) n * ld);
    double norm = 0.0;
    int pending = 0;
    run_cta(nthreads, [&](HostCta &cta) {
        // the sequence of eigs_schur_kernel (array_eigs.cu)
        for (int e = cta.tid; e < n * n; e += cta.nthreads) {
            H[(size_t) (e / n) * ld + e % n] = src[(size_t) e];
        }
        cta.sync();
        sstc::eigs_hessenberg(cta, H.data(), ld, V.data(), ort.data(), n, &sh);
        cta.sync();
        const int p = sstc::eigs_schur(cta, H.data(), ld, V.data(), val.data(), n, &sh, &norm, log.data(), log_cap,
                chunk.data(), chunk_cap);
        cta.sync();
        if (cta.tid == 0) {
            pending = p;
        }
        for (int e = cta.tid; e < n * n; e += cta.nthreads) {
            T[(size_t) e] = H[(size_t) (e / n) * ld + e % n];
        }
        cta.sync();
        if (p > 0) {  // V takes H's place (padded rows), the log is applied, V goes back
            for (int e = cta.tid; e < n * n; e += cta.nthreads) {
                Vs[(size_t) (e / n) * ld + e % n] = V[(size_t) e];
            }
            cta.sync();
            sstc::eigs_apply_rots(cta, Vs.data(), ld, n, log.data(), p, chunk.data(), chunk_cap);
            for (int e = cta.tid; e < n * n; e += cta.nthreads) {
                V[(size_t) e] = Vs[(size_t) (e / n) * ld + e % n];
            }
        }
    });
    if (pending < 0) {
        return 9;  // "Eigenvalue iteration did not converge"
    }
    if (norm == 0.0) {
        vec = V;  // LinearAlgebraOpsEigs.cpp:457-459
    } else {
        // one "thread" per eigenvalue, deliberately in ascending order (the reference goes descending): the columns
        // must be independent
        run_cta(nthreads, [&](HostCta &cta) {
            for (int c = cta.tid; c < n; c += cta.nthreads) {
                sstc::eigs_backsub_column(T.data(), X.data(), val.data(), n, c, norm);
            }
        });
        run_cta(nthreads, [&](HostCta &cta) {
            for (int e = cta.tid; e < n * n; e += cta.nthreads) {
                vec[(size_t) e] = sstc::eigs_backtransform_element(V.data(), X.data(), n, e / n, e % n);
            }
        });
    }
    f = fopen(argv[4], "wb");
    fwrite(vec.data(), 8, vec.size(), f);
    fwrite(val.data(), 8, val.size(), f);
    fclose(f);
    return 0;
}

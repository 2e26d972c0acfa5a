// TEST INFRASTRUCTURE ONLY -- drives shared_b200/csrc/eigs_core.cuh (the very code array_eigs.cu launches) on host
// threads. Usage: eigs_emu <n> <nthreads> <in.f64> <out.f64>; in = n*n doubles, out = n*n (vec) + 2n (val) doubles.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../shared_b200/csrc/eigs_core.cuh"
#include "cta_emu.h"

int main(int argc, char **argv) {
    if (argc != 5) {
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
    double norm = 0.0;
    run_cta(nthreads, [&](HostCta &cta) {
        for (int e = cta.tid; e < n * n; e += cta.nthreads) {
            H[(size_t) (e / n) * ld + e % n] = src[(size_t) e];
        }
        cta.sync();
        sstc::eigs_hessenberg(cta, H.data(), ld, V.data(), ort.data(), n, &sh);
        cta.sync();
        sstc::eigs_schur(cta, H.data(), ld, V.data(), val.data(), n, &sh, &norm);
        cta.sync();
        for (int e = cta.tid; e < n * n; e += cta.nthreads) {
            T[(size_t) e] = H[(size_t) (e / n) * ld + e % n];
        }
    });
    if (sh.mode == sstc::EIGS_FAILED) {
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

// TEST INFRASTRUCTURE ONLY -- the three phases of the executor's exclusive scan (sparse_core.cuh: scan_chunk_sum,
// scan_chunk_offsets, scan_chunk_write; launched by array_sparse.cu) on 256 host threads per "CTA". Usage:
// scan_emu <n> <seed>; exit 0 when the result equals the serial scan. Built with ThreadSanitizer.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../shared_b200/csrc/sparse_core.cuh"
#include "cta_emu.h"

int main(int argc, char **argv) {
    if (argc != 3) {
        return 2;
    }
    const int64_t n = atoll(argv[1]);
    unsigned seed = (unsigned) atoi(argv[2]);
    std::vector<int32_t> in((size_t) n + 1);
    for (int64_t i = 0; i < n; i++) {
        seed = seed * 1664525u + 1013904223u;
        in[(size_t) i] = (int32_t) ((seed >> 8) % 1000) + (i % 97 == 0 ? 2000000000 : 0);  // totals beyond int32
    }
    std::vector<int64_t> out((size_t) n + 1, -1), want((size_t) n + 1);
    int64_t run = 0;
    for (int64_t i = 0; i < n; i++) {
        want[(size_t) i] = run;
        run += in[(size_t) i];
    }
    want[(size_t) n] = run;
    const int64_t n_chunks = (n + sstc::kScanChunk - 1) / sstc::kScanChunk;
    std::vector<int64_t> sums((size_t) n_chunks + 1);
    const int grid = 3;  // fewer "CTAs" than chunks: the chunk loop of the kernels
    for (int b = 0; b < grid; b++) {
        std::vector<int64_t> smem(sstc::kScanThreads);
        run_cta(sstc::kScanThreads, [&](HostCta &cta) {
            for (int64_t c = b; c < n_chunks; c += grid) {
                sstc::scan_chunk_sum(cta, in.data(), n, c, sums.data(), smem.data());
            }
        });
    }
    {
        std::vector<int64_t> smem(sstc::kScanThreads);
        run_cta(sstc::kScanThreads, [&](HostCta &cta) {
            sstc::scan_chunk_offsets(cta, sums.data(), n_chunks, out.data() + n, smem.data());
        });
    }
    for (int b = 0; b < grid; b++) {
        std::vector<int64_t> smem(sstc::kScanThreads);
        run_cta(sstc::kScanThreads, [&](HostCta &cta) {
            for (int64_t c = b; c < n_chunks; c += grid) {
                sstc::scan_chunk_write(cta, in.data(), n, c, sums.data(), out.data(), smem.data());
            }
        });
    }
    for (int64_t i = 0; i <= n; i++) {
        if (out[(size_t) i] != want[(size_t) i]) {
            fprintf(stderr, "mismatch at %lld: %lld != %lld\n", (long long) i, (long long) out[(size_t) i],
                    (long long) want[(size_t) i]);
            return 1;
        }
    }
    return 0;
}

// TEST INFRASTRUCTURE ONLY -- runs a cooperative-thread-array body on host threads: tid / nthreads / sync() as the
// device code sees them, sync() being a pthread barrier. Built with -fsanitize=thread by tests/emu/Makefile so that a
// barrier missing from a kernel body in shared_b200/csrc/*_core.cuh is reported as a data race.
#pragma once
#include <pthread.h>

#include <functional>
#include <vector>

struct HostCta {
    int tid, nthreads;
    pthread_barrier_t *barrier;
    void sync() { pthread_barrier_wait(barrier); }
    void atomic_max(int *p, int v) {
        int cur = __atomic_load_n(p, __ATOMIC_RELAXED);
        while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
    }
};

inline void run_cta(int nthreads, const std::function<void(HostCta &)> &body) {
    pthread_barrier_t barrier;
    pthread_barrier_init(&barrier, nullptr, (unsigned) nthreads);
    struct Arg {
        HostCta cta;
        const std::function<void(HostCta &)> *body;
    };
    std::vector<Arg> args((size_t) nthreads);
    std::vector<pthread_t> th((size_t) nthreads);
    for (int t = 0; t < nthreads; t++) {
        args[(size_t) t].cta = HostCta{t, nthreads, &barrier};
        args[(size_t) t].body = &body;
        pthread_create(&th[(size_t) t], nullptr, [](void *p) -> void * {
            Arg *a = static_cast<Arg *>(p);
            (*a->body)(a->cta);
            return nullptr;
        }, &args[(size_t) t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[(size_t) t], nullptr);
    }
    pthread_barrier_destroy(&barrier);
}

// sstc_internal.h -- shared plumbing of libsst_cuda.so: error capture at the C ABI, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <stdexcept>
#include <string>

#include "../../include/sst_cuda.h"

namespace sstc {

// The reference reports failures as C++ exceptions turned into java.lang.RuntimeException at the JNI entry
// point (native/src/shared/Common.cpp:173-178). Here they stop at the C ABI: status code + message.
void set_last_error(const std::string &msg);
void clear_last_error();

extern std::atomic<int64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

inline void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    }
}

#define SSTC_CUDA(expr) ::sstc::cuda_check((expr), #expr)

inline void check_launch(const char *name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        throw std::runtime_error(std::string("kernel launch failed (") + name + "): " + cudaGetErrorString(e));
    }
    count_launch();
}

// Wraps the body of an extern "C" entry point.
template <typename F>
inline int guarded(F &&body) {
    try {
        clear_last_error();
        body();
        return 0;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return 1;
    } catch (...) {
        set_last_error("unknown error");
        return 1;
    }
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Grow-only device scratch owned by the host tier (one per thread so concurrent Java threads do not share).
struct Scratch {
    void *ptr = nullptr;
    int64_t bytes = 0;
    void *get(int64_t need);
    ~Scratch();
};

}  // namespace sstc

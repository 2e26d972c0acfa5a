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

// CUDA's current device is per host thread and starts at 0: a JVM worker thread that never called sstc_init would run
// on device 0 whatever the library was initialised with. Every entry point therefore binds a thread it sees for the
// first time to the device sstc_init selected (a thread that switches devices itself afterwards keeps its choice).
void bind_thread_device();
// Raises if a cross-GPU barrier kernel gave up waiting for a peer since the last check (sticky, see sstc_core.cu).
void check_barrier_error();
// Allocates the host-mapped error word for the current device (so that the first barrier launch does not).
void barrier_error_word_prepare();

// Wraps the body of an extern "C" entry point.
template <typename F>
inline int guarded(F &&body) {
    try {
        clear_last_error();
        bind_thread_device();
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

// The cross-GPU flag barrier (sstc_core.cu); epoch 0 = the kernel counts its own calls (graph-capturable).
void launch_peer_barrier(void *const *flag_arrays, int nranks, int rank, unsigned long long epoch, cudaStream_t stream);
// The barrier in two halves (a context uses either this pair or launch_peer_barrier, never both on one flag set).
void launch_peer_signal(void *const *flag_arrays, int nranks, int rank, cudaStream_t stream);
void launch_peer_wait(void *const *flag_arrays, int nranks, int rank, cudaStream_t stream);

// Grow-only device scratch owned by the host tier (one per thread so concurrent Java threads do not share).
struct Scratch {
    void *ptr = nullptr;
    int64_t bytes = 0;
    int device = -1;  // the device `ptr` lives on; a thread that moves to another device gets a fresh buffer there
    void *get(int64_t need);
    void release();
    ~Scratch();
};

// RAII device switch for code that drives several devices from one thread (the single-process slab transform).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        SSTC_CUDA(cudaGetDevice(&prev));
        if (prev != dev) {
            SSTC_CUDA(cudaSetDevice(dev));
        } else {
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

}  // namespace sstc

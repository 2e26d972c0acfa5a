// sstc_core.cu -- library lifecycle, error reporting and device-memory helpers of the C ABI.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "sstc_internal.h"

namespace sstc {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launch_count{0};

void set_last_error(const std::string &msg) { t_last_error = msg; }
void clear_last_error() { t_last_error.clear(); }

static std::atomic<int> g_device{-1};
static thread_local int t_bound_device = -1;
static std::atomic<bool> g_process_exiting{false};

void bind_thread_device() {
    const int dev = g_device.load(std::memory_order_relaxed);
    if (dev >= 0 && t_bound_device != dev) {
        SSTC_CUDA(cudaSetDevice(dev));
        t_bound_device = dev;
    }
}

void *Scratch::get(int64_t need) {
    int cur = 0;
    SSTC_CUDA(cudaGetDevice(&cur));
    if (ptr && device != cur) {
        release();
    }
    if (need > bytes) {
        release();
        SSTC_CUDA(cudaMalloc(&ptr, (size_t) need));
        bytes = need;
        device = cur;
    }
    return ptr;
}

void Scratch::release() {
    if (ptr) {
        int cur = -1;
        const bool moved = cudaGetDevice(&cur) == cudaSuccess && cur != device && device >= 0;
        if (moved) {
            cudaSetDevice(device);
        }
        cudaFree(ptr);  // errors ignored: at process exit the runtime may already be unloading
        if (moved) {
            cudaSetDevice(cur);
        }
        (void) cudaGetLastError();
    }
    ptr = nullptr;
    bytes = 0;
    device = -1;
}

Scratch::~Scratch() {
    // A worker thread that exits hands its staging buffers back (a JVM that churns threads would otherwise strand
    // them in HBM). During process teardown (static destructors, after exit() has begun) the CUDA runtime may already
    // be gone: leak then rather than call into it.
    if (!g_process_exiting.load()) {
        release();
    }
}

// Cross-GPU barrier for the one-process-per-GPU slab transform: every rank owns an array of 8 epoch slots that
// all ranks map (CUDA IPC). Rank r publishes `epoch` into slot r of every rank's array, then waits until its
// own array shows `epoch` from everyone. Launched after the kernel whose peer stores it fences, on the same
// stream, so those stores are complete before the flag is written. One kernel per rank, each on its own GPU,
// so the waiting ranks are always co-resident. The spin is bounded (~4 s) so that a missing peer cannot hang the
// device, and the barrier is sound or loud: a rank that gives up waiting records (epoch, missing rank) in a host-mapped
// word that every later entry point of the library checks (check_barrier_error), so the transform whose barrier failed
// -- or at the latest the next call -- raises instead of returning a slab that peers were still writing.
static unsigned long long *g_barrier_err_host = nullptr, *g_barrier_err_dev[16] = {nullptr};
static std::mutex g_barrier_err_mutex;

static unsigned long long *barrier_error_word() {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_barrier_err_mutex);
    if (!g_barrier_err_host) {
        SSTC_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&g_barrier_err_host), 64, cudaHostAllocMapped | cudaHostAllocPortable));
        *g_barrier_err_host = 0;
    }
    if (!g_barrier_err_dev[dev & 15]) {
        SSTC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&g_barrier_err_dev[dev & 15]), g_barrier_err_host, 0));
    }
    return g_barrier_err_dev[dev & 15];
}

void barrier_error_word_prepare() { (void) barrier_error_word(); }

void check_barrier_error() {
    if (g_barrier_err_host && *reinterpret_cast<volatile unsigned long long *>(g_barrier_err_host) != 0) {
        const unsigned long long w = *reinterpret_cast<volatile unsigned long long *>(g_barrier_err_host);
        *reinterpret_cast<volatile unsigned long long *>(g_barrier_err_host) = 0;
        throw std::runtime_error("peer barrier timed out: rank " + std::to_string((w >> 8) & 0xff) + " never saw rank " +
                std::to_string(w & 0xff) + " arrive (epoch " + std::to_string(w >> 16) + "); the slab it guards is invalid");
    }
}

struct PeerFlags {
    unsigned long long *ptr[8];
};

__global__ void peer_barrier_kernel(PeerFlags f, int nranks, int rank, unsigned long long epoch, unsigned long long *err) {
    const int r = threadIdx.x;
    if (epoch == 0) {
        // self-counting mode (CUDA-graph friendly: no per-launch argument changes): slot 8 of the rank's own
        // array counts this rank's barriers; every rank issues the same sequence, so the counts agree.
        __shared__ unsigned long long next;
        if (r == 0) {
            next = ++f.ptr[rank][8];
            f.ptr[rank][9] = next;  // the split barrier's wait counter stays in step (a context may switch styles)
        }
        __syncthreads();
        epoch = next;
    }
    if (r >= nranks) {
        return;
    }
    __threadfence_system();
    volatile unsigned long long *remote = f.ptr[r] + rank;
    *remote = epoch;
    __threadfence_system();
    volatile unsigned long long *local = f.ptr[rank] + r;
    const long long start = clock64();
    while (*local < epoch) {
        if (clock64() - start > 8000000000LL) {  // ~4 s: a peer died or never issued its barrier
            if (err) {
                *reinterpret_cast<volatile unsigned long long *>(err) = (epoch << 16) | ((unsigned long long) rank << 8) | (unsigned long long) r;
                __threadfence_system();
            }
            break;
        }
    }
    __threadfence_system();
}

// The same barrier in two halves, for schedules that keep the link busy: the SIGNAL goes out on the stream of the kernel
// whose stores it publishes, i.e. in front of the next exchange kernel's data (measured on 8 B200: a flag written while
// the next group's lines are already streaming queues behind tens of megabytes and arrives 50-80 us late); the WAIT
// runs on a side stream. Slot 8 of a rank's own array counts its signals, slot 9 its waits.
__global__ void peer_signal_kernel(PeerFlags f, int nranks, int rank) {
    __shared__ unsigned long long next;
    if (threadIdx.x == 0) {
        next = ++f.ptr[rank][8];
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r >= nranks) {
        return;
    }
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(f.ptr[r] + rank) = next;
}

__global__ void peer_wait_kernel(PeerFlags f, int nranks, int rank, unsigned long long *err) {
    __shared__ unsigned long long next;
    if (threadIdx.x == 0) {
        next = ++f.ptr[rank][9];
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r >= nranks) {
        return;
    }
    const unsigned long long epoch = next;
    volatile unsigned long long *local = f.ptr[rank] + r;
    const long long start = clock64();
    while (*local < epoch) {
        if (clock64() - start > 8000000000LL) {
            if (err) {
                *reinterpret_cast<volatile unsigned long long *>(err) = (epoch << 16) | ((unsigned long long) rank << 8) | (unsigned long long) r;
                __threadfence_system();
            }
            break;
        }
    }
    __threadfence_system();
}

static PeerFlags peer_flags(void *const *flag_arrays, int nranks) {
    PeerFlags f;
    for (int r = 0; r < 8; r++) {
        f.ptr[r] = r < nranks ? static_cast<unsigned long long *>(flag_arrays[r]) : nullptr;
    }
    return f;
}

void launch_peer_signal(void *const *flag_arrays, int nranks, int rank, cudaStream_t stream) {
    peer_signal_kernel<<<1, 32, 0, stream>>>(peer_flags(flag_arrays, nranks), nranks, rank);
    check_launch("peer_signal_kernel");
}

void launch_peer_wait(void *const *flag_arrays, int nranks, int rank, cudaStream_t stream) {
    peer_wait_kernel<<<1, 32, 0, stream>>>(peer_flags(flag_arrays, nranks), nranks, rank, barrier_error_word());
    check_launch("peer_wait_kernel");
}

void launch_peer_barrier(void *const *flag_arrays, int nranks, int rank, unsigned long long epoch, cudaStream_t stream) {
    PeerFlags f;
    for (int r = 0; r < 8; r++) {
        f.ptr[r] = r < nranks ? static_cast<unsigned long long *>(flag_arrays[r]) : nullptr;
    }
    peer_barrier_kernel<<<1, 32, 0, stream>>>(f, nranks, rank, epoch, barrier_error_word());
    check_launch("peer_barrier_kernel");
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_init(int device) {
    return guarded([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            throw std::runtime_error(
                    std::string("no CUDA device: libsst_cuda has no CPU fallback (") + cudaGetErrorString(e) + ")");
        }
        if (device < 0 || device >= n) {
            throw std::runtime_error("Invalid device");
        }
        SSTC_CUDA(cudaSetDevice(device));
        g_device.store(device);
        t_bound_device = device;
        static std::once_flag exit_hook;
        std::call_once(exit_hook, [] { atexit([] { g_process_exiting.store(true); }); });
        cudaDeviceProp prop;
        SSTC_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            throw std::runtime_error(std::string("libsst_cuda is built for sm_100a only; device is ") + prop.name);
        }
        SSTC_CUDA(cudaFree(0));
    });
}

int sstc_shutdown(void) {
    return guarded([&] { SSTC_CUDA(cudaDeviceSynchronize()); });
}

const char *sstc_last_error(void) { return t_last_error.c_str(); }

int sstc_abi_version(void) { return SSTC_ABI_VERSION; }

int64_t sstc_launch_count(void) { return g_launch_count.load(); }

int sstc_malloc(void **dptr, int64_t bytes) {
    return guarded([&] {
        if (!dptr || bytes < 0) {
            throw std::runtime_error("Invalid arguments");
        }
        *dptr = nullptr;
        if (bytes > 0) {
            SSTC_CUDA(cudaMalloc(dptr, (size_t) bytes));
        }
    });
}

int sstc_free(void *dptr) {
    return guarded([&] {
        if (dptr) {
            SSTC_CUDA(cudaFree(dptr));
        }
    });
}

int sstc_host_alloc(void **hptr, int64_t bytes) {
    return guarded([&] {
        if (!hptr || bytes < 0) {
            throw std::runtime_error("Invalid arguments");
        }
        *hptr = nullptr;
        if (bytes > 0) {
            SSTC_CUDA(cudaMallocHost(hptr, (size_t) bytes));
        }
    });
}

int sstc_host_free(void *hptr) {
    return guarded([&] {
        if (hptr) {
            SSTC_CUDA(cudaFreeHost(hptr));
        }
    });
}

int sstc_h2d(void *dst, const void *src, int64_t bytes, void *stream) {
    return guarded([&] {
        SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    });
}

int sstc_d2h(void *dst, const void *src, int64_t bytes, void *stream) {
    return guarded([&] {
        SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    });
}

int sstc_d2d(void *dst, const void *src, int64_t bytes, void *stream) {
    return guarded([&] {
        SSTC_CUDA(cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    });
}

int sstc_memset(void *dst, int value, int64_t bytes, void *stream) {
    return guarded([&] { SSTC_CUDA(cudaMemsetAsync(dst, value, (size_t) bytes, as_stream(stream))); });
}

int sstc_ipc_export(void *dptr, unsigned char handle[SSTC_IPC_HANDLE_BYTES]) {
    return guarded([&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == SSTC_IPC_HANDLE_BYTES, "handle size");
        if (!dptr || !handle) {
            throw std::runtime_error("Invalid arguments");
        }
        cudaIpcMemHandle_t h;
        SSTC_CUDA(cudaIpcGetMemHandle(&h, dptr));
        memcpy(handle, &h, sizeof(h));
    });
}

int sstc_ipc_open(const unsigned char handle[SSTC_IPC_HANDLE_BYTES], void **dptr) {
    return guarded([&] {
        if (!dptr || !handle) {
            throw std::runtime_error("Invalid arguments");
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof(h));
        SSTC_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    });
}

int sstc_ipc_close(void *dptr) {
    return guarded([&] {
        if (dptr) {
            SSTC_CUDA(cudaIpcCloseMemHandle(dptr));
        }
    });
}

int sstc_peer_barrier_dev(void *const *d_flag_arrays, int nranks, int rank, uint64_t epoch, void *stream) {
    return guarded([&] {
        if (!d_flag_arrays || nranks <= 0 || nranks > 8 || rank < 0 || rank >= nranks) {
            throw std::runtime_error("Invalid arguments");
        }
        check_barrier_error();
        launch_peer_barrier(d_flag_arrays, nranks, rank, (unsigned long long) epoch, as_stream(stream));
    });
}

int sstc_stream_sync(void *stream) {
    return guarded([&] {
        SSTC_CUDA(cudaStreamSynchronize(as_stream(stream)));
        check_barrier_error();
    });
}

}  // extern "C"

// ---- diagnostics: what a kernel can push into a peer's memory over NVLink (tools/exp_nvlink.py) -----------------
// mode 0: every thread loads and stores 16-byte pieces (8 in flight), the access pattern of the exchange pass;
// mode 1: the CTA pulls `piece` bytes into shared memory with one bulk copy (cp.async.bulk global -> shared, mbarrier)
//         and pushes them out with one bulk copy (shared -> global), two buffers in flight.
namespace sstc {

__global__ void __launch_bounds__(256) peer_copy_ldst_kernel(double2 *__restrict__ dst, const double2 *__restrict__ src, int64_t n) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        double2 v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            v[m] = src[i + m * stride];
        }
#pragma unroll
        for (int m = 0; m < 8; m++) {
            dst[i + m * stride] = v[m];
        }
    }
    for (; i < n; i += stride) {
        dst[i] = src[i];
    }
}

__global__ void __launch_bounds__(32) peer_copy_bulk_kernel(char *dst, const char *src, int64_t bytes, int piece) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    const int64_t npieces = bytes / piece;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) {
            const uint32_t ba = (uint32_t) __cvta_generic_to_shared(&bar[b]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t phase[2] = {0, 0};
        int64_t p = blockIdx.x;
        int buf = 0;
        // prologue: first load
        if (p < npieces) {
            const uint32_t ba = (uint32_t) __cvta_generic_to_shared(&bar[0]);
            const uint32_t sa = (uint32_t) __cvta_generic_to_shared(smem);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"((uint32_t) piece) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa),
                         "l"(src + p * piece), "r"((uint32_t) piece), "r"(ba)
                         : "memory");
        }
        for (; p < npieces; p += gridDim.x, buf ^= 1) {
            const int64_t next = p + gridDim.x;
            if (next < npieces) {
                // the other buffer's previous store must have read its source before it is overwritten
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                const uint32_t ba = (uint32_t) __cvta_generic_to_shared(&bar[buf ^ 1]);
                const uint32_t sa = (uint32_t) __cvta_generic_to_shared(smem + (size_t) (buf ^ 1) * piece);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"((uint32_t) piece) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa),
                             "l"(src + next * piece), "r"((uint32_t) piece), "r"(ba)
                             : "memory");
            }
            const uint32_t ba = (uint32_t) __cvta_generic_to_shared(&bar[buf]);
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done)
                             : "r"(ba), "r"(phase[buf])
                             : "memory");
            }
            phase[buf] ^= 1;
            const uint32_t sa = (uint32_t) __cvta_generic_to_shared(smem + (size_t) buf * piece);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + p * piece), "r"(sa),
                         "r"((uint32_t) piece)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

}  // namespace sstc

extern "C" int sstc_exp_peer_copy(void *dst, const void *src, int64_t bytes, int mode, int piece, int ctas, void *stream) {
    return sstc::guarded([&] {
        if (!dst || !src || bytes <= 0 || ctas <= 0 || (bytes & 15)) {
            throw std::runtime_error("Invalid arguments");
        }
        if (mode == 0) {
            sstc::peer_copy_ldst_kernel<<<ctas, 256, 0, sstc::as_stream(stream)>>>(static_cast<double2 *>(dst),
                    static_cast<const double2 *>(src), bytes / 16);
        } else {
            if (piece < 1024 || piece > 100 * 1024 || (piece & 127) || bytes % piece) {
                throw std::runtime_error("Invalid piece size");
            }
            SSTC_CUDA(cudaFuncSetAttribute(sstc::peer_copy_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * piece));
            sstc::peer_copy_bulk_kernel<<<ctas, 32, 2 * (size_t) piece, sstc::as_stream(stream)>>>(static_cast<char *>(dst),
                    static_cast<const char *>(src), bytes, piece);
        }
        sstc::check_launch("peer_copy_kernel");
    });
}

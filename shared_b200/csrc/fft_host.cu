// fft_host.cu -- HOST tier of the FftService half of the C ABI: sstc_fft on the caller's own arrays.
//
// Replaces Java_org_sharedx_fftw_Plan_transform (native/src/sharedx/BindingsX.cpp:26-29, Plan.cpp:44-99) and
// JavaFftService.{fft,ifft,rfft,rifft} (src/org/shared/fft/JavaFftService.java:53-94): the arrays are the JVM's
// double[] (pageable memory handed over by GetPrimitiveArrayCritical, native/src/shared/Common.cpp:72) or pinned
// buffers of a native caller.
//
// A transform through this tier is PCIe-bound (512^3 complex128: 2 GiB each way at ~55 GB/s = 2 x 39 ms against 2.2 ms of
// kernels), so the tier is built around the two copy engines:
//   * every call runs on its own "slot" (three non-blocking streams, device buffers, pinned staging rings, plans);
//     there is no global lock, so the D2H half of one caller's transform overlaps the H2D half of another's and both
//     PCIe directions stay busy (measured: 43 ms per 2 GiB + 2 GiB pair against 39 + 39 serial);
//   * within a call the input arrives in chunks of planes of the outermost dimension and every inner axis of a chunk
//     is transformed while the next chunk is still on the bus; the outermost axis is then transformed in groups of
//     columns and each group leaves for the host while the next is computed -- no kernel time is exposed;
//   * pageable arrays are staged through a ring of pinned buffers by a few copy threads, pipelined with the DMA
//     (the driver's own pageable path measured 11 GB/s H2D / 18 GB/s D2H on this pool; staged: the PCIe rate).
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "fft_internal.h"

namespace sstc {

// ---- options ------------------------------------------------------------------------------------------------

static std::atomic<int64_t> g_opt_slots{-1}, g_opt_threads{-1}, g_opt_pipeline{-1}, g_opt_chunk_mib{-1};

static int64_t option(std::atomic<int64_t> &o, const char *env, int64_t dflt) {
    const int64_t v = o.load();
    if (v >= 0) {
        return v;
    }
    const char *e = getenv(env);
    return e ? atoll(e) : dflt;
}

static int64_t opt_slots() { return std::max<int64_t>(1, option(g_opt_slots, "SSTC_HOST_SLOTS", 2)); }
static int64_t opt_pipeline() { return option(g_opt_pipeline, "SSTC_HOST_PIPELINE", 1); }
static size_t opt_chunk_bytes() {
    return (size_t) std::max<int64_t>(1, std::min<int64_t>(1024, option(g_opt_chunk_mib, "SSTC_HOST_CHUNK_MIB", 32))) << 20;
}
static int opt_threads() {
    const unsigned hw = std::thread::hardware_concurrency();
    const int64_t dflt = std::max(1, std::min(12, (int) (hw * 3 / 4)));  // measured on 16 cores: 8 -> 90 ms, 12 -> 83 ms
    return (int) std::max<int64_t>(1, std::min<int64_t>(64, option(g_opt_threads, "SSTC_HOST_COPY_THREADS", dflt)));
}

// ---- staging copies: a few persistent threads ------------------------------------------------------------------

// Staging copies stream: every byte is read once and written once and never touched again by the CPU, so the stores
// bypass the cache (no read-for-ownership of the destination line, no eviction of anything useful). Measured on this
// pool's hosts the staging copies, not the DMA, bound a pageable transform (44 GB/s with plain memcpy beside the DMA
// traffic against 55 GB/s on the bus); glibc only switches to non-temporal stores far above a task's 2 MiB.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stream_copy_avx2(char *d, const char *s, size_t n) {
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;
    if (head) {
        const size_t h = head < n ? head : n;
        memcpy(d, s, h);
        d += h;
        s += h;
        n -= h;
    }
    for (; n >= 128; d += 128, s += 128, n -= 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 96), e);
    }
    _mm_sfence();
    if (n) {
        memcpy(d, s, n);
    }
}
#endif

static void stream_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) {
        stream_copy_avx2(static_cast<char *>(dst), static_cast<const char *>(src), n);
        return;
    }
#endif
    memcpy(dst, src, n);
}

struct CopyTask {  // `rows` rows of `bytes` bytes each
    void *dst;
    const void *src;
    size_t bytes;
    size_t rows = 1, dpitch = 0, spitch = 0;
    void run() const {
        for (size_t r = 0; r < rows; r++) {
            stream_copy(static_cast<char *>(dst) + r * dpitch, static_cast<const char *>(src) + r * spitch, bytes);
        }
    }
};

class CopyPool {
  public:
    static CopyPool &get() {
        static CopyPool *pool = new CopyPool();  // never destroyed: worker threads may outlive static destruction
        return *pool;
    }

    // Runs the tasks to completion; the calling thread works too.
    void run(const std::vector<CopyTask> &tasks) {
        if (tasks.empty()) {
            return;
        }
        if (tasks.size() == 1) {
            tasks[0].run();
            return;
        }
        Batch batch;
        batch.remaining = (int) tasks.size();
        {
            std::lock_guard<std::mutex> lock(m_);
            ensure_workers();
            for (const CopyTask &t : tasks) {
                q_.push_back({t, &batch});
            }
        }
        cv_.notify_all();
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> lock(m_);
                if (q_.empty()) {
                    done_cv_.wait(lock, [&] { return batch.remaining == 0 || !q_.empty(); });
                    if (batch.remaining == 0) {
                        return;
                    }
                    continue;
                }
                it = q_.front();
                q_.pop_front();
            }
            execute(it);
        }
    }

  private:
    struct Batch {
        int remaining = 0;  // guarded by m_
    };
    struct Item {
        CopyTask t;
        Batch *b;
    };

    void execute(const Item &it) {
        it.t.run();
        std::lock_guard<std::mutex> lock(m_);
        if (--it.b->remaining == 0) {
            done_cv_.notify_all();
        }
    }

    void ensure_workers() {  // m_ held
        const int want = opt_threads() - 1;
        while ((int) workers_ < want) {
            workers_++;
            std::thread([this] {
                for (;;) {
                    Item it;
                    {
                        std::unique_lock<std::mutex> lock(m_);
                        cv_.wait(lock, [&] { return !q_.empty(); });
                        it = q_.front();
                        q_.pop_front();
                    }
                    execute(it);
                }
            }).detach();
        }
    }

    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    std::deque<Item> q_;
    int workers_ = 0;
};

static const size_t kCopyGrain = 2u << 20;

void host_copy(void *dst, const void *src, size_t bytes) {
    std::vector<CopyTask> tasks;
    for (size_t off = 0; off < bytes; off += kCopyGrain) {
        tasks.push_back(CopyTask{static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, std::min(kCopyGrain, bytes - off)});
    }
    CopyPool::get().run(tasks);
}

// rows x width bytes between pitched host arrays
void host_copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) {
    if (dpitch == width && spitch == width) {
        host_copy(dst, src, width * rows);
        return;
    }
    const size_t per = std::max<size_t>(1, kCopyGrain / std::max<size_t>(1, width));  // rows per task
    std::vector<CopyTask> tasks;
    for (size_t r = 0; r < rows; r += per) {
        CopyTask t{static_cast<char *>(dst) + r * dpitch, static_cast<const char *>(src) + r * spitch, width};
        t.rows = std::min(per, rows - r);
        t.dpitch = dpitch;
        t.spitch = spitch;
        tasks.push_back(t);
    }
    CopyPool::get().run(tasks);
}

// ---- slots ---------------------------------------------------------------------------------------------------

static const int kRing = 4;

struct HostSlot {
    int device = 0;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    Scratch d_in, d_out;
    std::vector<cudaEvent_t> events;
    size_t next_event = 0;
    void *ring[2][kRing] = {{nullptr}};  // [0]: host -> device staging, [1]: device -> host staging
    size_t ring_bytes = 0;
    cudaEvent_t ring_ev[2][kRing] = {{nullptr}};
    std::map<std::vector<int32_t>, sstc_fft_plan_s *> plans;

    cudaEvent_t event() {
        if (next_event == events.size()) {
            cudaEvent_t e;
            SSTC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            events.push_back(e);
        }
        return events[next_event++];
    }

    void need_ring(int dir, size_t bytes) {
        if (ring_bytes != bytes) {
            for (int d = 0; d < 2; d++) {
                for (int i = 0; i < kRing; i++) {
                    if (ring[d][i]) {
                        cudaFreeHost(ring[d][i]);
                        ring[d][i] = nullptr;
                    }
                }
            }
            ring_bytes = bytes;
        }
        for (int i = 0; i < kRing; i++) {
            if (!ring[dir][i]) {
                SSTC_CUDA(cudaHostAlloc(&ring[dir][i], bytes, cudaHostAllocDefault));
            }
            if (!ring_ev[dir][i]) {
                SSTC_CUDA(cudaEventCreateWithFlags(&ring_ev[dir][i], cudaEventDisableTiming));
            }
        }
    }

    void sync_all() {
        if (s_in) {
            cudaStreamSynchronize(s_in);
            cudaStreamSynchronize(s_run);
            cudaStreamSynchronize(s_out);
        }
    }
};

static std::mutex g_slots_mutex;
static std::condition_variable g_slots_cv;
static std::vector<HostSlot *> g_free_slots[16];
static int g_slots_created[16] = {0};

static HostSlot *acquire_slot() {
    int dev = 0;
    SSTC_CUDA(cudaGetDevice(&dev));
    const int d = dev & 15;
    std::unique_lock<std::mutex> lock(g_slots_mutex);
    for (;;) {
        if (!g_free_slots[d].empty()) {
            HostSlot *s = g_free_slots[d].back();
            g_free_slots[d].pop_back();
            return s;
        }
        if (g_slots_created[d] < opt_slots()) {
            g_slots_created[d]++;
            lock.unlock();
            HostSlot *s = new HostSlot();
            s->device = dev;
            try {
                int lo = 0, hi = 0;
                SSTC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                // non-blocking: other threads' legacy-stream work (the ArrayKernel host tier) must not serialise these
                SSTC_CUDA(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
                SSTC_CUDA(cudaStreamCreateWithFlags(&s->s_run, cudaStreamNonBlocking));
                SSTC_CUDA(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
            } catch (...) {
                for (cudaStream_t st : {s->s_in, s->s_run, s->s_out}) {
                    if (st) {
                        cudaStreamDestroy(st);
                    }
                }
                delete s;
                lock.lock();
                g_slots_created[d]--;
                g_slots_cv.notify_one();  // a caller waiting for a slot may now try to create one itself
                throw;
            }
            return s;
        }
        g_slots_cv.wait(lock);
    }
}

static void release_slot(HostSlot *s) {
    {
        std::lock_guard<std::mutex> lock(g_slots_mutex);
        g_free_slots[s->device & 15].push_back(s);
    }
    g_slots_cv.notify_one();
}

struct SlotLease {
    HostSlot *s;
    SlotLease() : s(acquire_slot()) { s->next_event = 0; }
    ~SlotLease() { release_slot(s); }
};

// ---- per-call statistics (sstc_host_last_stats): where a host-tier transform spends its time ------------------------

struct HostStats {
    double total_ms = 0, phase1_ms = 0, phase2_ms = 0;       // whole call; input side (H2D + inner axes); output side
    double stage_in_ms = 0, stage_out_ms = 0;                // host copies pageable -> ring, ring -> pageable
    double wait_in_ms = 0, wait_out_ms = 0;                  // host blocked on a staging buffer's DMA
    double pipelined = 0, in_pinned = 0, out_pinned = 0, chunks_in = 0, groups_out = 0;
};
static thread_local HostStats t_stats;

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- transfers -----------------------------------------------------------------------------------------------

bool host_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void) cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// One transform at a time per PCIe direction. Two calls that upload at the same time only halve each other's rate; what
// overlaps usefully is one call's upload with another's download. A call therefore holds the link's IN token while its
// input is on the bus and the OUT token while its output is, and the calls fall into step by themselves: caller A uploads,
// then downloads while caller B uploads, and so on -- both copy engines busy, neither direction shared.
static std::mutex g_link_in[16], g_link_out[16];
static std::atomic<int64_t> g_opt_link_tokens{-1};
static bool opt_link_tokens() { return option(g_opt_link_tokens, "SSTC_HOST_LINK_TOKENS", 1) != 0; }

static const size_t kSmallCopy = 1u << 20;  // below this the driver's own pageable path costs less than a staging hand-off

// One chunk host -> device on s_in. Returns the event recorded behind it.
static cudaEvent_t h2d_chunk(HostSlot &S, void *dst, const void *src, size_t bytes, bool pinned, int index) {
    if (pinned || bytes <= kSmallCopy) {
        SSTC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S.s_in));
        cudaEvent_t e = S.event();
        SSTC_CUDA(cudaEventRecord(e, S.s_in));
        return e;
    }
    const int j = index % kRing;
    double t0 = now_ms();
    if (index >= kRing) {
        SSTC_CUDA(cudaEventSynchronize(S.ring_ev[0][j]));  // the DMA that last read this staging buffer has finished
    }
    double t1 = now_ms();
    host_copy(S.ring[0][j], src, bytes);
    t_stats.wait_in_ms += t1 - t0;
    t_stats.stage_in_ms += now_ms() - t1;
    SSTC_CUDA(cudaMemcpyAsync(dst, S.ring[0][j], bytes, cudaMemcpyHostToDevice, S.s_in));
    SSTC_CUDA(cudaEventRecord(S.ring_ev[0][j], S.s_in));
    return S.ring_ev[0][j];
}

// Device -> host of a (rows x width) block with source pitch spitch into a destination with pitch dpitch, on s_out.
// Pageable destinations go through the ring; `Drain` finishes the host-side half (oldest first).
struct PendingOut {
    int j;
    void *dst;
    size_t dpitch, width, rows;
};

static void drain_one(HostSlot &S, std::deque<PendingOut> &pending) {
    PendingOut p = pending.front();
    pending.pop_front();
    double t0 = now_ms();
    SSTC_CUDA(cudaEventSynchronize(S.ring_ev[1][p.j]));
    double t1 = now_ms();
    host_copy_rows(p.dst, p.dpitch, S.ring[1][p.j], p.width, p.width, p.rows);
    t_stats.wait_out_ms += t1 - t0;
    t_stats.stage_out_ms += now_ms() - t1;
}

static void d2h_block(HostSlot &S, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
        bool pinned, int index, std::deque<PendingOut> &pending) {
    if (pinned || width * rows <= kSmallCopy) {
        if (rows == 1 || (dpitch == width && spitch == width)) {
            SSTC_CUDA(cudaMemcpyAsync(dst, src, width * rows, cudaMemcpyDeviceToHost, S.s_out));
        } else {
            SSTC_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, S.s_out));
        }
        return;
    }
    const int j = index % kRing;
    while ((int) pending.size() >= kRing) {
        drain_one(S, pending);  // frees the oldest staging buffer (== j, the ring is used in order)
    }
    if (rows == 1 || spitch == width) {
        SSTC_CUDA(cudaMemcpyAsync(S.ring[1][j], src, width * rows, cudaMemcpyDeviceToHost, S.s_out));
    } else {
        SSTC_CUDA(cudaMemcpy2DAsync(S.ring[1][j], width, src, spitch, width, rows, cudaMemcpyDeviceToHost, S.s_out));
    }
    SSTC_CUDA(cudaEventRecord(S.ring_ev[1][j], S.s_out));
    pending.push_back({j, dst, dpitch, width, rows});
}

// ---- schedules -----------------------------------------------------------------------------------------------

// Can the transform run as "inner axes chunk by chunk behind the H2D copy, outermost axis group by group in front of
// the D2H copy"? Needs the outermost axis to be the plan's last step, in place on the output, on a tile kernel.
static bool pipelinable(const sstc_fft_plan_s &p) {
    if (p.kind == SSTC_FFT_C2R || p.dims.size() < 2 || p.dims[0] < 2 || p.steps.size() < 2 || p.workspace_bytes != 0) {
        return false;
    }
    const Step &last = p.steps.back();
    if (last.axis_len != p.dims[0] || last.outer != 1 || last.mode != PASS_C2C || last.src != 1 || last.dst != 1 ||
            !tile_kernels_hold(last.axis_len)) {
        return false;
    }
    for (size_t k = 0; k + 1 < p.steps.size(); k++) {
        if (p.steps[k].outer % p.dims[0] != 0 || p.steps[k].dst != 1 || p.steps[k].src == 2) {
            return false;
        }
    }
    return true;
}

static void run_simple(HostSlot &S, sstc_fft_plan_s &p, const double *in, double *out, bool in_pinned, bool out_pinned) {
    const size_t in_bytes = (size_t) p.in_len * 8, out_bytes = (size_t) p.out_len * 8;
    char *d_in = static_cast<char *>(S.d_in.get((int64_t) in_bytes));
    char *d_out = static_cast<char *>(S.d_out.get((int64_t) out_bytes));
    const size_t chunk = opt_chunk_bytes();
    if ((!in_pinned && in_bytes > kSmallCopy) || (!out_pinned && out_bytes > kSmallCopy)) {
        S.need_ring(0, chunk);
        S.need_ring(1, chunk);
    }
    cudaEvent_t last = nullptr;
    int k = 0;
    for (size_t off = 0; off < in_bytes; off += chunk, k++) {
        last = h2d_chunk(S, d_in + off, reinterpret_cast<const char *>(in) + off, std::min(chunk, in_bytes - off), in_pinned, k);
    }
    if (last) {
        SSTC_CUDA(cudaStreamWaitEvent(S.s_run, last, 0));
    }
    execute_plan(p, reinterpret_cast<const double *>(d_in), reinterpret_cast<double *>(d_out), S.s_run);
    cudaEvent_t done = S.event();
    SSTC_CUDA(cudaEventRecord(done, S.s_run));
    SSTC_CUDA(cudaStreamWaitEvent(S.s_out, done, 0));
    std::deque<PendingOut> pending;
    k = 0;
    for (size_t off = 0; off < out_bytes; off += chunk, k++) {
        const size_t len = std::min(chunk, out_bytes - off);
        d2h_block(S, reinterpret_cast<char *>(out) + off, len, d_out + off, len, len, 1, out_pinned, k, pending);
    }
    while (!pending.empty()) {
        drain_one(S, pending);
    }
    SSTC_CUDA(cudaStreamSynchronize(S.s_out));
}

static void run_pipelined(HostSlot &S, sstc_fft_plan_s &p, const double *in, double *out, bool in_pinned, bool out_pinned) {
    const int64_t d0 = p.dims[0];
    const size_t in_bytes = (size_t) p.in_len * 8, out_bytes = (size_t) p.out_len * 8;
    const size_t in_plane = in_bytes / (size_t) d0, out_plane = out_bytes / (size_t) d0;  // bytes per plane of dimension 0
    char *d_in = static_cast<char *>(S.d_in.get((int64_t) in_bytes));
    char *d_out = static_cast<char *>(S.d_out.get((int64_t) out_bytes));
    const size_t chunk = opt_chunk_bytes();
    if (!in_pinned) {
        S.need_ring(0, chunk);
    }
    if (!out_pinned) {
        S.need_ring(1, chunk);
    }
    const size_t nsteps = p.steps.size();
    const bool tokens = opt_link_tokens();
    std::unique_lock<std::mutex> link_in(g_link_in[S.device & 15], std::defer_lock), link_out(g_link_out[S.device & 15], std::defer_lock);
    if (tokens) {
        link_in.lock();
    }
    const double t_begin = now_ms();
    cudaEvent_t last_arrived = nullptr;

    // phase 1: the input arrives in chunks; every inner axis of the planes that are complete is transformed behind it
    const int64_t min_planes = std::max<int64_t>(1, (int64_t) ((chunk / 2) / std::max<size_t>(1, in_plane)));
    int64_t planes_done = 0;
    int k = 0;
    for (size_t off = 0; off < in_bytes; k++) {
        const size_t len = std::min(chunk, in_bytes - off);
        cudaEvent_t arrived = h2d_chunk(S, d_in + off, reinterpret_cast<const char *>(in) + off, len, in_pinned, k);
        last_arrived = arrived;
        off += len;
        const int64_t ready = off == in_bytes ? d0 : (int64_t) (off / in_plane);
        if (ready > planes_done && (ready - planes_done >= min_planes || off == in_bytes)) {
            SSTC_CUDA(cudaStreamWaitEvent(S.s_run, arrived, 0));
            const int64_t np = ready - planes_done;
            for (size_t s = 0; s + 1 < nsteps; s++) {
                const Step &st = p.steps[s];
                const char *src = st.src == 0 ? d_in + (size_t) planes_done * in_plane : d_out + (size_t) planes_done * out_plane;
                char *dst = d_out + (size_t) planes_done * out_plane;
                fft_axis_pass(src, dst, st.outer / d0 * np, st.axis_len, st.inner, st.mode, p.inverse, st.scale, S.s_run);
            }
            planes_done = ready;
        }
    }

    if (tokens) {
        // the input has left the bus: hand the upload direction on, then wait for the download direction
        if (last_arrived) {
            SSTC_CUDA(cudaEventSynchronize(last_arrived));
        }
        link_in.unlock();
        link_out.lock();
    }
    t_stats.phase1_ms = now_ms() - t_begin;
    t_stats.chunks_in = k;
    const double t_phase2 = now_ms();
    // phase 2: the outermost axis, in groups of columns of the (d0, cols) view of the output; each group goes home
    // while the next one is transformed
    const Step &last = p.steps.back();
    const int64_t cols = (int64_t) (out_plane / 16);  // complex elements per plane
    int64_t gcols = (int64_t) (chunk / ((size_t) d0 * 16));
    gcols = std::max<int64_t>(8, gcols - gcols % 8);
    const int64_t layout[4] = {0, cols, 0, cols};
    std::deque<PendingOut> pending;
    k = 0;
    for (int64_t c0 = 0; c0 < cols; c0 += gcols, k++) {
        const int64_t nc = std::min(gcols, cols - c0);
        char *base = d_out + (size_t) c0 * 16;
        fft_axis_pass(base, base, 1, (int) d0, nc, PASS_C2C, p.inverse, last.scale, S.s_run, layout);
        cudaEvent_t done = S.event();
        SSTC_CUDA(cudaEventRecord(done, S.s_run));
        SSTC_CUDA(cudaStreamWaitEvent(S.s_out, done, 0));
        d2h_block(S, reinterpret_cast<char *>(out) + (size_t) c0 * 16, out_plane, base, out_plane, (size_t) nc * 16, (size_t) d0,
                out_pinned, k, pending);
    }
    while (!pending.empty()) {
        drain_one(S, pending);
    }
    SSTC_CUDA(cudaStreamSynchronize(S.s_out));
    t_stats.phase2_ms = now_ms() - t_phase2;
    t_stats.groups_out = k;
}

// ---- several devices: FftService.setHint("devices", "0,1,...") ---------------------------------------------------------
// With more than one device named, a 3-D c2c transform whose shape the slab transform accepts runs over all of them
// (sstc_fft3d_sharded_host): every device moves its slab over its own PCIe link.
static std::mutex g_devices_mutex;
static std::vector<int32_t> g_devices;
static std::map<std::vector<int32_t>, sstc_fft3d_sharded> g_sharded_plans;  // (devices..., dims...) -> plan

static bool sharded_shape(int kind, int rank, const int32_t *dims, int ndev) {
    if ((kind != SSTC_FFT_FORWARD && kind != SSTC_FFT_BACKWARD) || rank != 3 || ndev < 2) {
        return false;
    }
    const int d2 = dims[2];
    return dims[0] % ndev == 0 && dims[1] % ndev == 0 && d2 >= 8 && d2 <= 4096 && (d2 & (d2 - 1)) == 0 &&
            tile_kernels_hold(dims[0]) && tile_kernels_hold(dims[1]) &&
            (int64_t) dims[0] * dims[1] * d2 * 16 >= (int64_t) ndev * (8 << 20);
}

// true if the call was served by the multi-device plan
static bool run_on_devices(int kind, int rank, const int32_t *dims, const double *in, double *out) {
    std::lock_guard<std::mutex> lock(g_devices_mutex);  // one multi-device transform at a time: it owns every link
    const int ndev = (int) g_devices.size();
    if (!sharded_shape(kind, rank, dims, ndev)) {
        return false;
    }
    // Pageable arrays are bound by the host-side staging copies (44-75 GB/s for the whole box), not by the links: measured
    // on 4 GPUs, 92-112 ms over 2-4 devices against 86-92 ms on one (profiles/r2_exp_devices_hint.json). They stay on one.
    if (!host_is_pinned(in) || !host_is_pinned(out)) {
        return false;
    }
    std::vector<int32_t> key(g_devices);
    key.insert(key.end(), dims, dims + rank);
    auto it = g_sharded_plans.find(key);
    if (it == g_sharded_plans.end()) {
        sstc_fft3d_sharded plan = nullptr;
        if (sstc_fft3d_sharded_create(&plan, dims, ndev, g_devices.data()) != 0) {
            throw std::runtime_error(std::string(sstc_last_error()));
        }
        it = g_sharded_plans.emplace(key, plan).first;
    }
    if (sstc_fft3d_sharded_host(it->second, kind == SSTC_FFT_BACKWARD ? 1 : 0, in, out) != 0) {
        throw std::runtime_error(std::string(sstc_last_error()));
    }
    return true;
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_fft_set_devices(int ndev, const int32_t *devices) {
    return guarded([&] {
        if (ndev < 0 || ndev > 8 || (ndev > 0 && !devices)) {
            throw std::runtime_error("Invalid arguments");
        }
        int count = 0;
        SSTC_CUDA(cudaGetDeviceCount(&count));
        for (int i = 0; i < ndev; i++) {
            if (devices[i] < 0 || devices[i] >= count) {
                throw std::runtime_error("Invalid device");
            }
            for (int j = 0; j < i; j++) {
                if (devices[j] == devices[i]) {
                    throw std::runtime_error("Duplicate device");
                }
            }
        }
        std::lock_guard<std::mutex> lock(g_devices_mutex);
        for (auto &p : g_sharded_plans) {
            sstc_fft3d_sharded_destroy(p.second);
        }
        g_sharded_plans.clear();
        g_devices.assign(devices, devices + ndev);
    });
}

int sstc_fft(int kind, int rank, const int32_t *dims, const double *in, int64_t in_len, double *out, int64_t out_len) {
    return guarded([&] {
        if (!in || !out) {
            throw std::runtime_error("Invalid arguments");
        }
        int64_t a = 0, b = 0;
        transform_lengths(kind, rank, dims, a, b);
        if (in_len != a || out_len != b) {
            // Plan.cpp:88-90
            throw std::runtime_error("Input and/or output arrays do not have expected sizes");
        }
        if (run_on_devices(kind, rank, dims, in, out)) {
            return;
        }
        SlotLease lease;
        HostSlot &S = *lease.s;
        std::vector<int32_t> key;
        key.push_back(kind);
        key.insert(key.end(), dims, dims + rank);
        sstc_fft_plan_s *p = nullptr;
        auto it = S.plans.find(key);
        if (it == S.plans.end()) {
            p = new_plan(kind, rank, dims);
            S.plans[key] = p;
        } else {
            p = it->second;
        }
        const bool in_pinned = host_is_pinned(in), out_pinned = host_is_pinned(out);
        t_stats = HostStats();
        t_stats.in_pinned = in_pinned;
        t_stats.out_pinned = out_pinned;
        const double t_call = now_ms();
        try {
            if (opt_pipeline() && pipelinable(*p) && (size_t) (a + b) * 8 >= (8u << 20)) {
                t_stats.pipelined = 1;
                run_pipelined(S, *p, in, out, in_pinned, out_pinned);
            } else {
                run_simple(S, *p, in, out, in_pinned, out_pinned);
            }
            t_stats.total_ms = now_ms() - t_call;
        } catch (...) {
            S.sync_all();  // nothing of this call may still be in flight when the slot is handed on
            throw;
        }
    });
}

// Diagnostics / tests (no device needed): the staging copy the host tier uses for pageable arrays -- `rows` rows of `width`
// bytes between pitched host buffers, spread over the copy threads, non-temporal stores.
int sstc_exp_host_copy(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t rows) {
    try {
        clear_last_error();
        if (!dst || !src || width < 0 || rows < 0 || dpitch < width || spitch < width) {
            throw std::runtime_error("Invalid arguments");
        }
        host_copy_rows(dst, (size_t) dpitch, src, (size_t) spitch, (size_t) width, (size_t) rows);
        return 0;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return 1;
    }
}

int sstc_host_last_stats(double *out, int n) {
    return guarded([&] {
        if (!out || n < 0) {
            throw std::runtime_error("Invalid arguments");
        }
        const double v[12] = {t_stats.total_ms, t_stats.phase1_ms, t_stats.phase2_ms, t_stats.stage_in_ms, t_stats.stage_out_ms,
                t_stats.wait_in_ms, t_stats.wait_out_ms, t_stats.pipelined, t_stats.in_pinned, t_stats.out_pinned,
                t_stats.chunks_in, t_stats.groups_out};
        for (int i = 0; i < n && i < 12; i++) {
            out[i] = v[i];
        }
    });
}

int sstc_host_option(const char *name, int64_t value) {
    return guarded([&] {
        if (!name) {
            throw std::runtime_error("Invalid arguments");
        }
        const std::string n(name);
        if (n == "slots") {
            g_opt_slots.store(value);
        } else if (n == "copy_threads") {
            g_opt_threads.store(value);
        } else if (n == "pipeline") {
            g_opt_pipeline.store(value);
        } else if (n == "chunk_mib") {
            g_opt_chunk_mib.store(value);
        } else if (n == "link_tokens") {
            g_opt_link_tokens.store(value);
        } else {
            throw std::runtime_error("Unknown option");
        }
    });
}

}  // extern "C"

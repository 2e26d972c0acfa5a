// fft_slab.cu -- the slab-decomposed 3-D complex128 transform over the GPUs of one box, behind the C ABI.
//
// The reference has no distributed path; this is the one exchange step the north star adds.
// Layout contract for dims (d0, d1, d2) over P ranks (P divides d0 and d1, d2 a power of two the tile kernels hold):
//     natural slabs      rank g holds x[g*d0/P:(g+1)*d0/P, :, :]
//     transposed slabs   rank g holds X[:, g*d1/P:(g+1)*d1/P, :]      (like FFTW-MPI's TRANSPOSED_OUT)
// forward: natural -> transposed; inverse: transposed -> natural (scaled by 1/N).
//
// Schedule ("pipe"): the y pass runs locally; the z pass is the exchange pass -- it stores whole 16*d2-byte lines
// straight into the owners' receive buffers over NVLink, in G groups of y indices; after each group a flag barrier
// (peer_barrier_kernel) runs on a high-priority side stream and the x pass of that group's columns starts on a third
// stream beside the rest of the exchange. The inverse is the mirror image. A rank's schedule is recorded into a CUDA
// graph per (direction, input pointer, receive-buffer parity) after two eager runs.
//
// Two front ends share one per-rank context (sstc_slab):
//   * one process per GPU (bench.py under torchrun, shared_b200/sharded.py): the caller maps the peers' buffers
//     (sstc_ipc_*) and passes the pointers to sstc_slab_create;
//   * one process, several devices (sstc_fft3d_sharded_*, what a JVM provider binds): the library allocates the
//     buffers on every device and enables peer access itself.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <algorithm>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "fft_internal.h"

struct sstc_slab_s {
    int device = 0, nranks = 1, rank = 0;
    int d0 = 0, d1 = 0, d2 = 0, x_loc = 0, y_loc = 0;
    int64_t n_local = 0, block = 0, n_total = 0;
    void *recv[2][8];
    void *flags[8];
    double *work = nullptr;
    cudaStream_t s_main = nullptr, s_z = nullptr, s_bar = nullptr, s_x = nullptr;
    std::vector<cudaEvent_t> events;
    size_t next_event = 0;
    int groups = 4, chunks = 1, exchange_ctas = 0, bulk = 0, use_graph = 1;
    int order = 0, head_ctas = 0;  // see issue_forward
    int split_barrier = 0;         // see barrier_after_exchange
    // forward_begin (the local y pass of a LATER transform issued ahead, beside the exchange of the current one)
    double *work_alt = nullptr;                 // second work buffer: begun transforms alternate between the two
    cudaStream_t s_y = nullptr;
    cudaEvent_t ev_begin_enter[2] = {nullptr, nullptr}, ev_y_done[2] = {nullptr, nullptr}, ev_step_done = nullptr;
    bool have_step_done = false;
    uint64_t begin_count = 0;
    struct Pending {
        const double *in;
        int widx;
    };
    std::deque<Pending> pending;
    uint64_t step = 0;
    std::map<std::tuple<int, const void *, int>, cudaGraphExec_t> graphs;
    std::map<std::tuple<int, const void *, int>, int64_t> graph_kernels;
    std::map<std::pair<int, const void *>, int> seen;
    // diagnostics ("trace" option): timing events behind every launch of an eagerly issued step
    int trace = 0;
    std::vector<cudaEvent_t> marks;
    std::vector<std::string> mark_labels;
    size_t next_mark = 0;
};

namespace sstc {

static cudaEvent_t slab_event(sstc_slab_s &S) {
    if (S.next_event == S.events.size()) {
        cudaEvent_t e;
        SSTC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        S.events.push_back(e);
    }
    return S.events[S.next_event++];
}

static cudaEvent_t record_on(sstc_slab_s &S, cudaStream_t s) {
    cudaEvent_t e = slab_event(S);
    SSTC_CUDA(cudaEventRecord(e, s));
    return e;
}

// trace: a timing event on stream s, labelled
static void mark(sstc_slab_s &S, cudaStream_t s, const std::string &label) {
    if (!S.trace) {
        return;
    }
    if (S.next_mark == S.marks.size()) {
        cudaEvent_t e;
        SSTC_CUDA(cudaEventCreate(&e));
        S.marks.push_back(e);
        S.mark_labels.push_back(label);
    }
    S.mark_labels[S.next_mark] = label;
    SSTC_CUDA(cudaEventRecord(S.marks[S.next_mark++], s));
}

static int pick_divisor(int want, int of) {
    for (int g : {want, 4, 2, 1}) {
        if (g >= 1 && of % g == 0) {
            return g;
        }
    }
    return 1;
}

// Cross-rank barrier behind the exchange kernel just issued on s_z; the event it returns fires when every rank has
// passed it. Default: one kernel on the high-priority side stream publishes this rank's flag and waits for the peers'.
// split_barrier = 1: the SIGNAL is issued on s_z itself, so that it leaves ahead of the next exchange kernel's data (a flag
// written while the next group's lines are already streaming queues behind them and arrives 50-80 us late: the step trace,
// profiles/r2_sharded_8gpu_sweeps_and_traces.json), and only the WAIT runs on the side stream. The x passes then start
// earlier, but the step does not get shorter (0.525 against 0.514 ms at 8 GPUs: the extra kernel between the exchange
// launches costs what the earlier start gains, because the passes that run beside the exchange slow it down either way).
static cudaEvent_t barrier_after_exchange(sstc_slab_s &S) {
    if (S.split_barrier) {
        launch_peer_signal(S.flags, S.nranks, S.rank, S.s_z);
        SSTC_CUDA(cudaStreamWaitEvent(S.s_bar, record_on(S, S.s_z), 0));
        launch_peer_wait(S.flags, S.nranks, S.rank, S.s_bar);
    } else {
        SSTC_CUDA(cudaStreamWaitEvent(S.s_bar, record_on(S, S.s_z), 0));
        launch_peer_barrier(S.flags, S.nranks, S.rank, 0, S.s_bar);
    }
    return record_on(S, S.s_bar);
}

// The forward schedule on S.s_main (and the side streams, forked from and joined back into it).
//
// Tasks: Y(c) = the local y pass of plane chunk c; E(c, g) = the exchange (z pass + line scatter) of y group g of chunk
// c, which needs Y(c); X(g) = the x pass of the received columns of group g, which needs E(c, g) of EVERY chunk on every
// rank (one barrier). The link is the bottleneck, so the order of the E tasks decides what is exposed:
//   order 0  E(*, 0) chunk by chunk behind the y pass, then E(all, g) for g = 1..: every x pass but the last runs beside
//            the exchange, but the exchange of group 0 crawls behind the y pass (small launches that share the SMs);
//   order 1  chunk-major first, group-major last: E(c, all groups) as ONE launch per chunk for c < C-1, uncapped (nothing
//            needs the SMs yet but the next chunk's y pass), then the last chunk group by group with the x passes beside
//            it. The link starts after 1/C of the y pass and never waits for a small launch.
// x == nullptr: the y pass of this transform was issued ahead (sstc_slab_forward_begin) into `work_buf` and s_main has
// already waited for it; the exchange then starts at full speed, group 0 included.
static void issue_forward(sstc_slab_s &S, const double *x, int parity, double *work_buf) {
    const int P = S.nranks;
    const int G = pick_divisor(S.groups, S.y_loc), C = x ? pick_divisor(S.chunks, S.x_loc) : 1;
    const int yg = S.y_loc / G, xc = S.x_loc / C;
    const int64_t cols = (int64_t) yg * S.d2, kstride = (int64_t) S.y_loc * S.d2, plane = (int64_t) S.d1 * S.d2;
    char *work = reinterpret_cast<char *>(work_buf);
    void *base[8];
    for (int q = 0; q < P; q++) {
        base[q] = static_cast<char *>(S.recv[parity][q]) + (size_t) S.rank * S.block * 16;
    }
    auto chunk_dst = [&](int c, void **dst) {
        for (int q = 0; q < P; q++) {
            dst[q] = static_cast<char *>(base[q]) + (size_t) c * xc * kstride * 16;
        }
    };
    // y pass, local, in C chunks of planes
    std::vector<cudaEvent_t> y_done;
    mark(S, S.s_main, "start");
    for (int c = 0; c < C; c++) {
        const size_t off = (size_t) c * xc * plane * 16;
        if (x) {
            fft_axis_pass(reinterpret_cast<const char *>(x) + off, work + off, xc, S.d1, S.d2, PASS_C2C, 0, 1.0, S.s_main);
        }
        y_done.push_back(record_on(S, S.s_main));  // (pre-begun: the fork point of the side streams)
        mark(S, S.s_main, "y" + std::to_string(c) + " done");
    }
    const bool two_phase = S.order == 1 && C > 1;
    if (two_phase) {
        for (int c = 0; c + 1 < C; c++) {
            SSTC_CUDA(cudaStreamWaitEvent(S.s_z, y_done[(size_t) c], 0));
            void *dst[8];
            chunk_dst(c, dst);
            mark(S, S.s_z, "E" + std::to_string(c) + ".* begin");
            fft_lines_scatter(reinterpret_cast<const double *>(work + (size_t) c * xc * plane * 16), dst, P, xc, S.d1, S.d2, 0,
                    S.y_loc, 0, 1.0, S.head_ctas, S.bulk, S.s_z);
            mark(S, S.s_z, "E" + std::to_string(c) + ".* done");
        }
    }
    cudaEvent_t joined = nullptr;
    for (int g = 0; g < G; g++) {
        if (two_phase) {
            const int c = C - 1;
            if (g == 0) {
                SSTC_CUDA(cudaStreamWaitEvent(S.s_z, y_done[(size_t) c], 0));
            }
            void *dst[8];
            chunk_dst(c, dst);
            mark(S, S.s_z, "e" + std::to_string(g) + " begin");
            fft_lines_scatter(reinterpret_cast<const double *>(work + (size_t) c * xc * plane * 16), dst, P, xc, S.d1, S.d2, g * yg,
                    yg, 0, 1.0, S.exchange_ctas, S.bulk, S.s_z);
            mark(S, S.s_z, "e" + std::to_string(g) + " done");
        } else if (g == 0) {
            for (int c = 0; c < C; c++) {
                SSTC_CUDA(cudaStreamWaitEvent(S.s_z, y_done[(size_t) c], 0));
                void *dst[8];
                chunk_dst(c, dst);
                mark(S, S.s_z, "e0." + std::to_string(c) + " begin");
                fft_lines_scatter(reinterpret_cast<const double *>(work + (size_t) c * xc * plane * 16), dst, P, xc, S.d1, S.d2, 0,
                        yg, 0, 1.0, S.exchange_ctas, S.bulk, S.s_z);
                mark(S, S.s_z, "e0." + std::to_string(c) + " done");
            }
        } else {
            mark(S, S.s_z, "e" + std::to_string(g) + " begin");
            fft_lines_scatter(work_buf, base, P, S.x_loc, S.d1, S.d2, g * yg, yg, 0, 1.0, S.exchange_ctas, S.bulk, S.s_z);
            mark(S, S.s_z, "e" + std::to_string(g) + " done");
        }
        cudaEvent_t arrived = barrier_after_exchange(S);
        mark(S, S.s_bar, "barrier" + std::to_string(g) + " passed");
        SSTC_CUDA(cudaStreamWaitEvent(S.s_x, arrived, 0));
        char *ptr = static_cast<char *>(S.recv[parity][S.rank]) + (size_t) g * cols * 16;
        const int64_t layout[4] = {0, kstride, 0, kstride};
        mark(S, S.s_x, "x" + std::to_string(g) + " begin");
        fft_axis_pass(ptr, ptr, 1, S.d0, cols, PASS_C2C, 0, 1.0, S.s_x, layout);  // x on this group's columns
        mark(S, S.s_x, "x" + std::to_string(g) + " done");
        joined = record_on(S, S.s_x);
    }
    SSTC_CUDA(cudaStreamWaitEvent(S.s_main, joined, 0));
    mark(S, S.s_main, "end");
}

// The inverse schedule: x pass locally, the z pass scatters whole lines into the owners' natural slabs plane group by
// plane group, the y pass (with the 1/N scale) follows group by group.
static void issue_inverse(sstc_slab_s &S, const double *y, int parity) {
    const int P = S.nranks;
    const int G = pick_divisor(S.groups, S.x_loc);
    const int xg = S.x_loc / G;
    const int64_t plane = (int64_t) S.d1 * S.d2;
    fft_axis_pass(y, S.work, 1, S.d0, (int64_t) S.y_loc * S.d2, PASS_C2C, 1, 1.0, S.s_main);
    SSTC_CUDA(cudaStreamWaitEvent(S.s_z, record_on(S, S.s_main), 0));
    void *base[8];
    for (int q = 0; q < P; q++) {
        base[q] = static_cast<char *>(S.recv[parity][q]) + (size_t) S.rank * S.y_loc * S.d2 * 16;
    }
    cudaEvent_t joined = nullptr;
    for (int g = 0; g < G; g++) {
        fft_planes_scatter(S.work, base, P, S.x_loc, S.y_loc, S.d1, S.d2, g * xg, xg, 1, 1.0, S.exchange_ctas, S.bulk, S.s_z);
        cudaEvent_t arrived = barrier_after_exchange(S);
        SSTC_CUDA(cudaStreamWaitEvent(S.s_x, arrived, 0));
        char *ptr = static_cast<char *>(S.recv[parity][S.rank]) + (size_t) g * xg * plane * 16;
        fft_axis_pass(ptr, ptr, xg, S.d1, S.d2, PASS_C2C, 1, 1.0 / (double) S.n_total, S.s_x);
        joined = record_on(S, S.s_x);
    }
    SSTC_CUDA(cudaStreamWaitEvent(S.s_main, joined, 0));
}

static void slab_run(sstc_slab_s &S, int dir, const double *in, double **out, cudaStream_t user) {
    if (!in || !out) {
        throw std::runtime_error("Invalid arguments");
    }
    check_barrier_error();
    DeviceGuard guard(S.device);
    const int parity = (int) (S.step & 1);
    S.step++;
    S.next_event = 0;
    S.next_mark = 0;
    // the caller's stream joins in front of and behind the schedule, which lives on the context's own streams
    cudaEvent_t enter = slab_event(S);
    SSTC_CUDA(cudaEventRecord(enter, user));
    SSTC_CUDA(cudaStreamWaitEvent(S.s_main, enter, 0));
    // a transform whose y pass was issued ahead?
    double *work_buf = S.work;
    bool begun = false;
    int gdir = dir, gpar = parity;
    if (dir == 0 && !S.pending.empty() && S.pending.front().in == in) {
        const int widx = S.pending.front().widx;
        S.pending.pop_front();
        begun = true;
        work_buf = widx ? S.work_alt : S.work;
        SSTC_CUDA(cudaStreamWaitEvent(S.s_main, S.ev_y_done[widx], 0));
        gdir = 2;
        gpar = parity * 2 + widx;
    } else if (!S.pending.empty()) {
        throw std::runtime_error("a begun forward transform is pending: finish it with sstc_slab_forward on the same input first");
    }
    const auto key = std::make_tuple(gdir, (const void *) in, gpar);
    auto it = S.graphs.find(key);
    if (it == S.graphs.end() && S.use_graph && !S.trace && S.seen[std::make_pair(gdir, (const void *) in)] >= 2) {
        // both parities have run eagerly with this input (attributes set, tables built): record this one
        cudaGraph_t graph = nullptr;
        const int64_t before = g_launch_count.load();
        SSTC_CUDA(cudaStreamBeginCapture(S.s_main, cudaStreamCaptureModeRelaxed));
        try {
            if (dir == 0) {
                issue_forward(S, begun ? nullptr : in, parity, work_buf);
            } else {
                issue_inverse(S, in, parity);
            }
        } catch (...) {
            cudaStreamEndCapture(S.s_main, &graph);
            if (graph) {
                cudaGraphDestroy(graph);
            }
            throw;
        }
        SSTC_CUDA(cudaStreamEndCapture(S.s_main, &graph));
        const int64_t kernels = g_launch_count.load() - before;
        g_launch_count.fetch_sub(kernels);  // recorded, not launched: replays are counted below
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        SSTC_CUDA(e);
        S.graphs[key] = exec;
        S.graph_kernels[key] = kernels;
        it = S.graphs.find(key);
    }
    if (it != S.graphs.end()) {
        SSTC_CUDA(cudaGraphLaunch(it->second, S.s_main));
        count_launch((int) S.graph_kernels[key]);
    } else {
        S.seen[std::make_pair(gdir, (const void *) in)]++;
        if (dir == 0) {
            issue_forward(S, begun ? nullptr : in, parity, work_buf);
        } else {
            issue_inverse(S, in, parity);
        }
    }
    if (!S.ev_step_done) {
        SSTC_CUDA(cudaEventCreateWithFlags(&S.ev_step_done, cudaEventDisableTiming));
    }
    SSTC_CUDA(cudaEventRecord(S.ev_step_done, S.s_main));  // what a y pass issued ahead waits for before it reuses a work buffer
    S.have_step_done = true;
    cudaEvent_t leave = slab_event(S);
    SSTC_CUDA(cudaEventRecord(leave, S.s_main));
    SSTC_CUDA(cudaStreamWaitEvent(user, leave, 0));
    *out = static_cast<double *>(S.recv[parity][S.rank]);
}

// The local y pass of a transform that sstc_slab_forward(in) will finish later, issued NOW on its own stream so that it
// runs beside the exchange of the transform in flight: in a stream of independent transforms the head of every step (the
// pass that must precede the first byte on the link) hides behind its predecessor. At most two transforms can be begun
// ahead (two work buffers); they are finished in the order they were begun.
static void slab_begin(sstc_slab_s &S, const double *in, cudaStream_t user) {
    if (!in) {
        throw std::runtime_error("Invalid arguments");
    }
    if (S.pending.size() >= 2) {
        throw std::runtime_error("two forward transforms are already begun");
    }
    DeviceGuard guard(S.device);
    if (!S.s_y) {
        SSTC_CUDA(cudaStreamCreateWithFlags(&S.s_y, cudaStreamNonBlocking));
        SSTC_CUDA(cudaMalloc(reinterpret_cast<void **>(&S.work_alt), (size_t) S.n_local * 16));
        for (int i = 0; i < 2; i++) {
            SSTC_CUDA(cudaEventCreateWithFlags(&S.ev_begin_enter[i], cudaEventDisableTiming));
            SSTC_CUDA(cudaEventCreateWithFlags(&S.ev_y_done[i], cudaEventDisableTiming));
        }
    }
    const int widx = (int) (S.begin_count++ & 1);
    SSTC_CUDA(cudaEventRecord(S.ev_begin_enter[widx], user));
    SSTC_CUDA(cudaStreamWaitEvent(S.s_y, S.ev_begin_enter[widx], 0));
    if (S.have_step_done) {
        // this work buffer was last read two steps ago; the latest finished step is a (more than) sufficient fence
        SSTC_CUDA(cudaStreamWaitEvent(S.s_y, S.ev_step_done, 0));
    }
    fft_axis_pass(in, widx ? S.work_alt : S.work, S.x_loc, S.d1, S.d2, PASS_C2C, 0, 1.0, S.s_y);
    SSTC_CUDA(cudaEventRecord(S.ev_y_done[widx], S.s_y));
    S.pending.push_back({in, widx});
}

static void slab_destroy(sstc_slab_s *S) {
    if (!S) {
        return;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(S->device);
    for (cudaStream_t s : {S->s_main, S->s_z, S->s_bar, S->s_x, S->s_y}) {
        if (s) {
            cudaStreamSynchronize(s);
        }
    }
    for (cudaEvent_t e : {S->ev_begin_enter[0], S->ev_begin_enter[1], S->ev_y_done[0], S->ev_y_done[1], S->ev_step_done}) {
        if (e) {
            cudaEventDestroy(e);
        }
    }
    if (S->s_y) {
        cudaStreamDestroy(S->s_y);
    }
    if (S->work_alt) {
        cudaFree(S->work_alt);
    }
    for (auto &g : S->graphs) {
        cudaGraphExecDestroy(g.second);
    }
    for (cudaEvent_t e : S->events) {
        cudaEventDestroy(e);
    }
    for (cudaStream_t s : {S->s_main, S->s_z, S->s_bar, S->s_x}) {
        if (s) {
            cudaStreamDestroy(s);
        }
    }
    if (S->work) {
        cudaFree(S->work);
    }
    if (prev >= 0) {
        cudaSetDevice(prev);
    }
    (void) cudaGetLastError();
    delete S;
}

static sstc_slab_s *slab_create(const int32_t *dims, int nranks, int rank, void *const *recv0, void *const *recv1,
        void *const *flags) {
    if (!dims || !recv0 || !recv1 || !flags || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) {
        throw std::runtime_error("Invalid arguments");
    }
    for (int i = 0; i < 3; i++) {
        if (dims[i] <= 0) {
            throw std::runtime_error("Invalid dimensions");
        }
    }
    if (dims[0] % nranks || dims[1] % nranks) {
        throw std::runtime_error("the rank count must divide d0 and d1");
    }
    const int d2 = dims[2];
    if (!(d2 >= 8 && d2 <= 4096 && (d2 & (d2 - 1)) == 0)) {
        throw std::runtime_error("the pipelined slab exchange needs a power-of-two d2 between 8 and 4096");
    }
    if (!tile_kernels_hold(dims[0]) || !tile_kernels_hold(dims[1])) {
        throw std::runtime_error("the slab transform needs d0 and d1 within the tile kernels' range");
    }
    std::unique_ptr<sstc_slab_s, void (*)(sstc_slab_s *)> S(new sstc_slab_s(), slab_destroy);
    SSTC_CUDA(cudaGetDevice(&S->device));
    S->nranks = nranks;
    S->rank = rank;
    S->d0 = dims[0];
    S->d1 = dims[1];
    S->d2 = d2;
    S->x_loc = S->d0 / nranks;
    S->y_loc = S->d1 / nranks;
    S->n_local = (int64_t) S->x_loc * S->d1 * S->d2;
    S->block = (int64_t) S->x_loc * S->y_loc * S->d2;
    S->n_total = (int64_t) S->d0 * S->d1 * S->d2;
    for (int q = 0; q < 8; q++) {
        S->recv[0][q] = q < nranks ? recv0[q] : nullptr;
        S->recv[1][q] = q < nranks ? recv1[q] : nullptr;
        S->flags[q] = q < nranks ? flags[q] : nullptr;
        if (q < nranks && (!recv0[q] || !recv1[q] || !flags[q])) {
            throw std::runtime_error("Invalid arguments");
        }
    }
    SSTC_CUDA(cudaMalloc(reinterpret_cast<void **>(&S->work), (size_t) S->n_local * 16));
    int lo = 0, hi = 0;
    SSTC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SSTC_CUDA(cudaStreamCreateWithFlags(&S->s_main, cudaStreamNonBlocking));
    SSTC_CUDA(cudaStreamCreateWithFlags(&S->s_z, cudaStreamNonBlocking));
    // the barrier must not queue behind exchange CTAs, and the x pass must win SM slots from the pending exchange CTAs
    SSTC_CUDA(cudaStreamCreateWithPriority(&S->s_bar, cudaStreamNonBlocking, hi));
    SSTC_CUDA(cudaStreamCreateWithPriority(&S->s_x, cudaStreamNonBlocking, hi));
    cudaDeviceProp prop;
    SSTC_CUDA(cudaGetDeviceProperties(&prop, S->device));
    // measured on B200s at 512^3 (profiles/r1_sweep_pipe_2_4_8gpu.json): one exchange CTA per SM keeps NVLink busy at 4
    // and 8 ranks and leaves room for the x pass beside it; at 2 ranks the exchange moves half of the data and wants two
    S->exchange_ctas = prop.multiProcessorCount * (nranks == 2 ? 2 : 1);
    S->groups = 4;
    S->chunks = nranks >= 8 ? 4 : 1;
    // Everything a first launch allocates or configures (twiddle tables, shared-memory attributes, the barrier's error
    // word) happens here, on scratch data: a call that synchronises the device must never sit between two ranks'
    // enqueues when one thread drives several devices, or a spinning barrier kernel would wait for work that is not
    // yet submitted.
    fft_axis_pass(S->work, S->work, 1, S->d1, S->d2, PASS_C2C, 0, 1.0, S->s_main);
    if (nranks > 1) {
        void *self[8];
        for (int q = 0; q < nranks; q++) {
            self[q] = S->work;  // all destinations local: same kernel instantiation, no peer traffic
        }
        fft_lines_scatter(S->work, self, nranks, 1, S->d1, S->d2, 0, 1, 0, 1.0, 1, 0, S->s_main);
        fft_lines_scatter(S->work, self, nranks, 1, S->d1, S->d2, 0, 1, 0, 1.0, 1, 1, S->s_main);
        fft_lines_scatter(S->work, self, nranks, 1, S->d1, S->d2, 0, 1, 0, 1.0, 1, 2, S->s_main);
    }
    const int64_t layout[4] = {0, 8, 0, 8};
    fft_axis_pass(S->work, S->work, 1, S->d0, 8, PASS_C2C, 0, 1.0, S->s_main, layout);
    fft_axis_pass(S->work, S->work, 1, S->d0, 8, PASS_C2C, 1, 1.0, S->s_main);
    barrier_error_word_prepare();
    SSTC_CUDA(cudaStreamSynchronize(S->s_main));
    return S.release();
}

}  // namespace sstc

using namespace sstc;

// ---- single process, several devices -----------------------------------------------------------------------------

struct sstc_fft3d_sharded_s {
    int ndev = 0;
    int32_t dims[3] = {0, 0, 0};
    std::vector<int> devices;
    std::vector<sstc_slab_s *> ranks;
    std::vector<void *> recv[2], flags;   // per device, library-owned
    std::vector<cudaStream_t> streams;    // per device: the stream the host-tier front end drives
    std::vector<double *> in_buf;         // per device: staging for the host-tier front end (grow on first use)
    std::vector<void *> pinned[2];        // host staging for pageable arrays, per device, [0] in / [1] out
    size_t pinned_bytes = 0;
};

namespace sstc {

static void sharded_destroy(sstc_fft3d_sharded_s *H) {
    if (!H) {
        return;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (sstc_slab_s *r : H->ranks) {
        slab_destroy(r);
    }
    for (int i = 0; i < H->ndev; i++) {
        cudaSetDevice(H->devices[(size_t) i]);
        for (int par = 0; par < 2; par++) {
            if ((size_t) i < H->recv[par].size() && H->recv[par][(size_t) i]) {
                cudaFree(H->recv[par][(size_t) i]);
            }
            if ((size_t) i < H->pinned[par].size() && H->pinned[par][(size_t) i]) {
                cudaFreeHost(H->pinned[par][(size_t) i]);
            }
        }
        if ((size_t) i < H->flags.size() && H->flags[(size_t) i]) {
            cudaFree(H->flags[(size_t) i]);
        }
        if ((size_t) i < H->in_buf.size() && H->in_buf[(size_t) i]) {
            cudaFree(H->in_buf[(size_t) i]);
        }
        if ((size_t) i < H->streams.size() && H->streams[(size_t) i]) {
            cudaStreamDestroy(H->streams[(size_t) i]);
        }
    }
    if (prev >= 0) {
        cudaSetDevice(prev);
    }
    (void) cudaGetLastError();
    delete H;
}

static void sharded_sync(sstc_fft3d_sharded_s &H) {
    for (int i = 0; i < H.ndev; i++) {
        DeviceGuard g(H.devices[(size_t) i]);
        SSTC_CUDA(cudaStreamSynchronize(H.ranks[(size_t) i]->s_main));
        SSTC_CUDA(cudaStreamSynchronize(H.streams[(size_t) i]));
    }
    check_barrier_error();
}

}  // namespace sstc

extern "C" {

int sstc_slab_create(sstc_slab *slab, const int32_t *dims, int nranks, int rank, void *const *d_recv0, void *const *d_recv1,
        void *const *d_flags) {
    return guarded([&] {
        if (!slab) {
            throw std::runtime_error("Invalid arguments");
        }
        *slab = nullptr;
        *slab = slab_create(dims, nranks, rank, d_recv0, d_recv1, d_flags);
    });
}

int sstc_slab_forward(sstc_slab slab, const double *d_in, double **d_out, void *stream) {
    return guarded([&] {
        if (!slab) {
            throw std::runtime_error("Invalid arguments");
        }
        slab_run(*slab, 0, d_in, d_out, as_stream(stream));
    });
}

int sstc_slab_forward_begin(sstc_slab slab, const double *d_in, void *stream) {
    return guarded([&] {
        if (!slab) {
            throw std::runtime_error("Invalid arguments");
        }
        check_barrier_error();
        slab_begin(*slab, d_in, as_stream(stream));
    });
}

int sstc_slab_inverse(sstc_slab slab, const double *d_in, double **d_out, void *stream) {
    return guarded([&] {
        if (!slab) {
            throw std::runtime_error("Invalid arguments");
        }
        slab_run(*slab, 1, d_in, d_out, as_stream(stream));
    });
}

int sstc_slab_set_option(sstc_slab slab, const char *name, int64_t value) {
    return guarded([&] {
        if (!slab || !name) {
            throw std::runtime_error("Invalid arguments");
        }
        const std::string n(name);
        if (value < 0) {
            throw std::runtime_error("Invalid option value");
        }
        if (n == "groups") {
            slab->groups = (int) std::max<int64_t>(1, value);
        } else if (n == "chunks") {
            slab->chunks = (int) std::max<int64_t>(1, value);
        } else if (n == "exchange_ctas") {
            slab->exchange_ctas = (int) value;
        } else if (n == "bulk_store") {
            if (value > 2) {
                throw std::runtime_error("Invalid option value");
            }
            slab->bulk = (int) value;
        } else if (n == "order") {
            slab->order = value ? 1 : 0;
        } else if (n == "split_barrier") {
            // both styles count their calls in the rank's own flag slots: switch only while no step is in flight, and on
            // every rank alike
            slab->split_barrier = value ? 1 : 0;
        } else if (n == "head_ctas") {
            slab->head_ctas = (int) value;
        } else if (n == "graph") {
            slab->use_graph = value ? 1 : 0;
        } else if (n == "trace") {
            slab->trace = value ? 1 : 0;
        } else {
            throw std::runtime_error("Unknown option");
        }
        // recorded schedules were built with the old values
        DeviceGuard guard(slab->device);
        for (cudaStream_t s : {slab->s_main, slab->s_z, slab->s_bar, slab->s_x}) {
            SSTC_CUDA(cudaStreamSynchronize(s));
        }
        for (auto &g : slab->graphs) {
            cudaGraphExecDestroy(g.second);
        }
        slab->graphs.clear();
        slab->graph_kernels.clear();
    });
}

// Diagnostics: after a step issued with the "trace" option (and a synchronise), writes up to `max` records of
// "label<TAB>milliseconds since the step's start<NEWLINE>" into buf; returns the number of bytes written.
int64_t sstc_slab_trace(sstc_slab slab, char *buf, int64_t max) {
    int64_t used = 0;
    guarded([&] {
        if (!slab || !buf || slab->next_mark == 0) {
            return;
        }
        DeviceGuard guard(slab->device);
        std::string out;
        for (size_t i = 0; i < slab->next_mark; i++) {
            SSTC_CUDA(cudaEventSynchronize(slab->marks[i]));
            float ms = 0;
            SSTC_CUDA(cudaEventElapsedTime(&ms, slab->marks[0], slab->marks[i]));
            out += slab->mark_labels[i] + "\t" + std::to_string(ms) + "\n";
        }
        used = std::min<int64_t>(max, (int64_t) out.size());
        memcpy(buf, out.data(), (size_t) used);
    });
    return used;
}

int64_t sstc_slab_graph_count(sstc_slab slab) { return slab ? (int64_t) slab->graphs.size() : 0; }

int sstc_slab_destroy(sstc_slab slab) {
    return guarded([&] { slab_destroy(slab); });
}

int sstc_fft3d_sharded_create(sstc_fft3d_sharded *plan, const int32_t *dims, int ndev, const int32_t *devices) {
    return guarded([&] {
        if (!plan || !dims || ndev < 1 || ndev > 8) {
            throw std::runtime_error("Invalid arguments");
        }
        *plan = nullptr;
        int count = 0;
        SSTC_CUDA(cudaGetDeviceCount(&count));
        std::unique_ptr<sstc_fft3d_sharded_s, void (*)(sstc_fft3d_sharded_s *)> H(new sstc_fft3d_sharded_s(), sharded_destroy);
        H->ndev = ndev;
        for (int i = 0; i < 3; i++) {
            H->dims[i] = dims[i];
        }
        for (int i = 0; i < ndev; i++) {
            const int d = devices ? devices[i] : i;
            if (d < 0 || d >= count) {
                throw std::runtime_error("Invalid device");
            }
            for (int j = 0; j < i; j++) {
                if (H->devices[(size_t) j] == d) {
                    throw std::runtime_error("Duplicate device");
                }
            }
            H->devices.push_back(d);
        }
        if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || dims[0] % ndev || dims[1] % ndev) {
            throw std::runtime_error("the device count must divide d0 and d1");
        }
        const size_t n_local = (size_t) dims[0] / ndev * dims[1] * dims[2];
        for (int par = 0; par < 2; par++) {
            H->recv[par].assign((size_t) ndev, nullptr);
            H->pinned[par].assign((size_t) ndev, nullptr);
        }
        H->flags.assign((size_t) ndev, nullptr);
        H->streams.assign((size_t) ndev, nullptr);
        H->in_buf.assign((size_t) ndev, nullptr);
        for (int i = 0; i < ndev; i++) {
            DeviceGuard g(H->devices[(size_t) i]);
            for (int j = 0; j < ndev; j++) {
                if (j != i) {
                    int can = 0;
                    SSTC_CUDA(cudaDeviceCanAccessPeer(&can, H->devices[(size_t) i], H->devices[(size_t) j]));
                    if (!can) {
                        throw std::runtime_error("devices cannot access each other's memory (no NVLink / P2P path)");
                    }
                    cudaError_t e = cudaDeviceEnablePeerAccess(H->devices[(size_t) j], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                        SSTC_CUDA(e);
                    }
                    (void) cudaGetLastError();
                }
            }
            SSTC_CUDA(cudaMalloc(&H->recv[0][(size_t) i], n_local * 16));
            SSTC_CUDA(cudaMalloc(&H->recv[1][(size_t) i], n_local * 16));
            SSTC_CUDA(cudaMalloc(&H->flags[(size_t) i], 128));
            SSTC_CUDA(cudaMemset(H->flags[(size_t) i], 0, 128));
            SSTC_CUDA(cudaStreamCreateWithFlags(&H->streams[(size_t) i], cudaStreamNonBlocking));
            SSTC_CUDA(cudaDeviceSynchronize());
        }
        for (int i = 0; i < ndev; i++) {
            DeviceGuard g(H->devices[(size_t) i]);
            H->ranks.push_back(slab_create(dims, ndev, i, H->recv[0].data(), H->recv[1].data(), H->flags.data()));
        }
        *plan = H.release();
    });
}

static void sharded_run(sstc_fft3d_sharded_s &H, int dir, const double *const *d_in, double **d_out, void *const *streams) {
    if (!d_in || !d_out) {
        throw std::runtime_error("Invalid arguments");
    }
    // every rank's schedule is enqueued before any of them is waited for: the barrier kernels of one device spin until
    // the other devices' kernels, enqueued a few microseconds later by this same thread, publish their flags
    for (int i = 0; i < H.ndev; i++) {
        DeviceGuard g(H.devices[(size_t) i]);
        cudaStream_t s = streams ? as_stream(streams[i]) : H.streams[(size_t) i];
        slab_run(*H.ranks[(size_t) i], dir, d_in[i], &d_out[i], s);
    }
}

int sstc_fft3d_sharded_forward(sstc_fft3d_sharded plan, const double *const *d_in, double **d_out, void *const *streams) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        sharded_run(*plan, 0, d_in, d_out, streams);
    });
}

int sstc_fft3d_sharded_inverse(sstc_fft3d_sharded plan, const double *const *d_in, double **d_out, void *const *streams) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        sharded_run(*plan, 1, d_in, d_out, streams);
    });
}

int sstc_fft3d_sharded_sync(sstc_fft3d_sharded plan) {
    return guarded([&] {
        if (!plan) {
            throw std::runtime_error("Invalid arguments");
        }
        sharded_sync(*plan);
    });
}

int sstc_fft3d_sharded_set_option(sstc_fft3d_sharded plan, const char *name, int64_t value) {
    if (!plan) {
        return guarded([] { throw std::runtime_error("Invalid arguments"); });
    }
    for (sstc_slab_s *r : plan->ranks) {
        const int rc = sstc_slab_set_option(r, name, value);
        if (rc != 0) {
            return rc;
        }
    }
    return 0;
}

// HOST tier over several devices: `in` / `out` are whole (d0, d1, d2) complex arrays in host memory, natural order on
// both sides. Every device moves its own slab over its own PCIe link, which is what makes this faster than the
// one-device host tier beyond the kernels: 2 GiB each way is 39 ms on one link.
int sstc_fft3d_sharded_host(sstc_fft3d_sharded plan, int inverse, const double *in, double *out) {
    return guarded([&] {
        if (!plan || !in || !out) {
            throw std::runtime_error("Invalid arguments");
        }
        sstc_fft3d_sharded_s &H = *plan;
        const int P = H.ndev;
        const size_t d0 = (size_t) H.dims[0], d1 = (size_t) H.dims[1], d2 = (size_t) H.dims[2];
        const size_t n_local = d0 / P * d1 * d2, slab_bytes = n_local * 16;
        const size_t row_t = d1 / P * d2 * 16;  // bytes of one x row of a transposed slab
        const size_t row_n = d1 * d2 * 16;      // bytes of one x row of the whole array
        const bool in_pinned = host_is_pinned(in), out_pinned = host_is_pinned(out);
        if ((!in_pinned || !out_pinned) && H.pinned_bytes != slab_bytes) {
            for (int i = 0; i < P; i++) {
                DeviceGuard g(H.devices[(size_t) i]);
                for (int par = 0; par < 2; par++) {
                    if (H.pinned[par][(size_t) i]) {
                        cudaFreeHost(H.pinned[par][(size_t) i]);
                        H.pinned[par][(size_t) i] = nullptr;
                    }
                    SSTC_CUDA(cudaHostAlloc(&H.pinned[par][(size_t) i], slab_bytes, cudaHostAllocPortable));
                }
            }
            H.pinned_bytes = slab_bytes;
        }
        std::vector<const double *> d_in((size_t) P);
        std::vector<double *> d_out((size_t) P);
        for (int i = 0; i < P; i++) {
            DeviceGuard g(H.devices[(size_t) i]);
            if (!H.in_buf[(size_t) i]) {
                SSTC_CUDA(cudaMalloc(reinterpret_cast<void **>(&H.in_buf[(size_t) i]), slab_bytes));
            }
            cudaStream_t s = H.streams[(size_t) i];
            const char *src = reinterpret_cast<const char *>(in);
            if (!inverse) {  // natural slab i: a contiguous range of the array
                const char *from = src + (size_t) i * slab_bytes;
                if (!in_pinned) {
                    host_copy(H.pinned[0][(size_t) i], from, slab_bytes);
                    from = static_cast<const char *>(H.pinned[0][(size_t) i]);
                }
                SSTC_CUDA(cudaMemcpyAsync(H.in_buf[(size_t) i], from, slab_bytes, cudaMemcpyHostToDevice, s));
            } else {  // transposed slab i: d0 rows of row_t bytes, row_n apart in the array
                const char *from = src + (size_t) i * row_t;
                if (!in_pinned) {
                    char *stage = static_cast<char *>(H.pinned[0][(size_t) i]);
                    host_copy_rows(stage, row_t, from, row_n, row_t, d0);
                    SSTC_CUDA(cudaMemcpyAsync(H.in_buf[(size_t) i], stage, slab_bytes, cudaMemcpyHostToDevice, s));
                } else {
                    SSTC_CUDA(cudaMemcpy2DAsync(H.in_buf[(size_t) i], row_t, from, row_n, row_t, d0, cudaMemcpyHostToDevice, s));
                }
            }
            d_in[(size_t) i] = H.in_buf[(size_t) i];
        }
        sharded_run(H, inverse ? 1 : 0, d_in.data(), d_out.data(), nullptr);
        for (int i = 0; i < P; i++) {
            DeviceGuard g(H.devices[(size_t) i]);
            cudaStream_t s = H.streams[(size_t) i];
            char *dst = reinterpret_cast<char *>(out);
            if (!inverse) {  // transposed slab i -> columns [i*row_t, (i+1)*row_t) of every x row
                if (out_pinned) {
                    SSTC_CUDA(cudaMemcpy2DAsync(dst + (size_t) i * row_t, row_n, d_out[(size_t) i], row_t, row_t, d0,
                            cudaMemcpyDeviceToHost, s));
                } else {
                    SSTC_CUDA(cudaMemcpyAsync(H.pinned[1][(size_t) i], d_out[(size_t) i], slab_bytes, cudaMemcpyDeviceToHost, s));
                }
            } else {
                void *to = out_pinned ? (void *) (dst + (size_t) i * slab_bytes) : H.pinned[1][(size_t) i];
                SSTC_CUDA(cudaMemcpyAsync(to, d_out[(size_t) i], slab_bytes, cudaMemcpyDeviceToHost, s));
            }
        }
        for (int i = 0; i < P; i++) {
            DeviceGuard g(H.devices[(size_t) i]);
            SSTC_CUDA(cudaStreamSynchronize(H.streams[(size_t) i]));
            if (!out_pinned) {
                char *dst = reinterpret_cast<char *>(out);
                const char *stage = static_cast<const char *>(H.pinned[1][(size_t) i]);
                if (!inverse) {
                    host_copy_rows(dst + (size_t) i * row_t, row_n, stage, row_t, row_t, d0);
                } else {
                    host_copy(dst + (size_t) i * slab_bytes, stage, slab_bytes);
                }
            }
        }
        check_barrier_error();
    });
}

int sstc_fft3d_sharded_destroy(sstc_fft3d_sharded plan) {
    return guarded([&] { sharded_destroy(plan); });
}

}  // extern "C"

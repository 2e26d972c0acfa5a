/*
 * sst_cuda.h -- C ABI of libsst_cuda.so, the B200 (sm_100a) provider for the FftService and ArrayKernel
 * service-provider interfaces of carsomyr/Shared.
 *
 * This is the drop-in boundary: the JNI glue (shared_b200/csrc/jni_glue.cpp), the Python host mirror
 * (shared_b200/) and bench.py all bind exactly these entry points. Plain pointers and sizes only.
 * Citations are relative to /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; sstc_last_error() then holds the message
 *     that the JNI glue raises as java.lang.RuntimeException (native/src/shared/Common.cpp:173-178 does the
 *     same for the reference's C++ exceptions). No C++ exception crosses this ABI.
 *   - HOST tier (`sstc_<op>`): pointers are host memory, exactly the arrays the Java SPI passes; the library
 *     stages them through HBM (H2D, kernels, D2H) and returns when the result is in the caller's array.
 *   - DEVICE tier (`sstc_<op>_dev`, plans): pointers are device memory on the current device and the call
 *     is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *   - complex data is interleaved re/im float64, row-major, as in org.shared.array.ComplexArray.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef SST_CUDA_H
#define SST_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSTC_ABI_VERSION 2

/* ---- library / device ------------------------------------------------------------------------------- */

/* Replaces JNI_OnLoad's provider registration (native/src/shared_common/Library.cpp:55-77): selects the
 * CUDA device the calling thread will use and checks it is an sm_100 part. */
int sstc_init(int device);
int sstc_shutdown(void);
/* Thread-local message of the last failed call ("" if none). */
const char *sstc_last_error(void);
int sstc_abi_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t sstc_launch_count(void);

/* Device-tier memory helpers (cudaMalloc / cudaFree / cudaMemcpy on the current device). */
int sstc_malloc(void **dptr, int64_t bytes);
int sstc_free(void *dptr);
int sstc_host_alloc(void **hptr, int64_t bytes); /* pinned host memory */
int sstc_host_free(void *hptr);
int sstc_h2d(void *dst, const void *src, int64_t bytes, void *stream);
int sstc_d2h(void *dst, const void *src, int64_t bytes, void *stream);
int sstc_d2d(void *dst, const void *src, int64_t bytes, void *stream);
int sstc_memset(void *dst, int value, int64_t bytes, void *stream);
int sstc_stream_sync(void *stream);

/* ---- FftService ------------------------------------------------------------------------------------ */

/* Transform kinds; values are those of org.sharedx.fftw.Plan (src_extensions/org/sharedx/fftw/Plan.java:73-88). */
#define SSTC_FFT_R2C 0      /* FftService.rfft  (src/org/shared/fft/FftService.java:50)  */
#define SSTC_FFT_C2R 1      /* FftService.rifft (FftService.java:62), scaled by 1/N     */
#define SSTC_FFT_FORWARD 2  /* FftService.fft   (FftService.java:74), e^{-2 pi i jk/n}, unscaled */
#define SSTC_FFT_BACKWARD 3 /* FftService.ifft  (FftService.java:86), scaled by 1/N     */

/* Expected array lengths in doubles for a transform; the rules of Plan::getTransformParameters
 * (native/src/sharedx/Plan.cpp:230-320). `dims` are the logical dims (real dims for R2C/C2R, no trailing 2). */
int sstc_fft_lengths(int kind, int rank, const int32_t *dims, int64_t *in_len, int64_t *out_len);

/* HOST tier. Replaces Java_org_sharedx_fftw_Plan_transform (native/src/sharedx/BindingsX.cpp:26-29,
 * Plan.cpp:44-99) and JavaFftService.{fft,ifft,rfft,rifft} (src/org/shared/fft/JavaFftService.java:53-94).
 * `in` is preserved. Length mismatch fails with "Input and/or output arrays do not have expected sizes".
 * Re-entrant: concurrent callers run on separate slots (streams, buffers) and share the two PCIe directions. The
 * arrays may be pageable (a JVM heap array under GetPrimitiveArrayCritical) or pinned; pageable ones are staged through
 * a pinned ring by copy threads, pipelined with the DMA. The input is transferred in chunks of planes of dims[0] and
 * every inner axis is transformed behind the copy; dims[0] is transformed in column groups in front of the D2H copy. */
int sstc_fft(int kind, int rank, const int32_t *dims, const double *in, int64_t in_len, double *out,
        int64_t out_len);

/* FftService.setHint("devices", "0,1,2,3") on the C side: with two or more devices named, host-tier 3-D c2c transforms
 * whose shape the slab transform accepts (device count divides d0 and d1, d2 a power of two in [8, 4096], at least 8 MiB
 * per device) run over all of them through sstc_fft3d_sharded_host; everything else stays on the calling thread's
 * device. ndev <= 1 returns to one device. Errors: "Invalid device", "Duplicate device". */
int sstc_fft_set_devices(int ndev, const int32_t *devices);

/* Tunables of the host tier (process-wide; the FftService.setHint analogue for the transfer side). Names:
 *   "slots"        transforms that may be in flight at once per device (default 2; each holds its own device buffers,
 *                  streams and pinned staging ring, so that one caller's D2H overlaps another's H2D);
 *   "copy_threads" threads that move pageable arrays through the pinned staging ring (default min(12, 3/4 of the cores));
 *   "chunk_mib"    staging / pipelining granularity (default 32);
 *   "pipeline"     0 = H2D, all passes, D2H strictly one after the other (for A/B measurements), 1 = default;
 *   "link_tokens"  1 (default) = one call at a time per PCIe direction, so that concurrent callers fall into
 *                  upload-while-the-other-downloads step instead of sharing a direction; 0 = no arbitration.
 * Environment: SSTC_HOST_SLOTS, SSTC_HOST_COPY_THREADS, SSTC_HOST_CHUNK_MIB, SSTC_HOST_PIPELINE, SSTC_HOST_LINK_TOKENS. */
int sstc_host_option(const char *name, int64_t value);
/* The staging copy of the host tier on its own (tests, diagnostics; needs no device): `rows` rows of `width` bytes from a
 * host buffer with pitch `spitch` into one with pitch `dpitch`, spread over the copy threads, non-temporal stores. */
int sstc_exp_host_copy(void *dst, int64_t dpitch, const void *src, int64_t spitch, int64_t width, int64_t rows);

/* Where the calling thread's last sstc_fft spent its time (milliseconds): out[0..11] = total, input phase (H2D + inner
 * axes), output phase (outermost axis + D2H), host copies pageable -> ring, ring -> pageable, host blocked on a ring
 * buffer's DMA (in, out), pipelined?, in pinned?, out pinned?, input chunks, output groups. */
int sstc_host_last_stats(double *out, int n);

/* DEVICE tier. A plan replaces the fftw_plan held by org.sharedx.fftw.Plan (Plan.cpp:101-141 create,
 * :143-175 destroy); creation is serialised internally, execution is re-entrant per stream. */
typedef struct sstc_fft_plan_s *sstc_fft_plan;
int sstc_fft_plan_create(sstc_fft_plan *plan, int kind, int rank, const int32_t *dims);
int sstc_fft_plan_execute(sstc_fft_plan plan, const double *d_in, double *d_out, void *stream);
int sstc_fft_plan_destroy(sstc_fft_plan plan);
/* Bytes of device workspace the plan owns, and the number of kernel launches one execute issues. */
int64_t sstc_fft_plan_workspace(sstc_fft_plan plan);
int sstc_fft_plan_launches(sstc_fft_plan plan);
/* Plan tunables (FftService.setHint analogue; FftwService.java:93-143). Known names: "transposed_passes"
 * (c2c, rank >= 3: the two outermost passes exchange their data through a transposed workspace so that
 * neither loads with the outermost stride; -1 = automatic, 0 = never, 1 = whenever legal). Unknown names
 * fail with "Unknown hint" (JavaFftService.java:96-104). Not thread-safe against a concurrent execute. */
int sstc_fft_plan_set_hint(sstc_fft_plan plan, const char *name, int64_t value);

/* Fused FFT-domain correlation on device-resident data -- ConvolutionCache.convolve(ComplexArray cIm, RealArray ker)
 * (src/org/shared/fft/ConvolutionCache.java:99-114): d_out = rifft(d_spec .* d_kernel_spec)[0:out_dims[0], ...],
 * a dense row-major real array of dims out_dims (out_dims[d] = dims[d] - kernelSize[d] + 1 <= dims[d]). `plan`
 * is a C2R plan for the image dims; d_spec is the image's half-spectrum (what rfft returned), d_kernel_spec the
 * cached conj(rfft(padded kernel)) (ConvolutionCache.createCacheable, :117-138; FftCache.java:83-105 keeps it).
 * The reference runs eMul, rifft and subarray as three whole-array operations through JVM arrays; here the
 * product is formed while the first inverse pass loads and the crop happens while the last pass stores, so the
 * product, the full-size inverse and the index tables of subarray never exist. Neither input is modified. */
int sstc_fft_convolve_dev(sstc_fft_plan plan, const double *d_spec, const double *d_kernel_spec, double *d_out,
        const int32_t *out_dims, void *stream);

/* One axis pass of a row-major complex array viewed as (outer, len, inner): the building block the
 * slab-sharded multi-GPU transform (shared_b200/sharded.py) composes around its exchange step.
 * `inverse` selects e^{+2 pi i jk/n}; every output is multiplied by `scale`. In place when d_in == d_out. */
int sstc_fft_axis_dev(const double *d_in, double *d_out, int64_t outer, int32_t len, int64_t inner, int inverse,
        double scale, void *stream);

/* The same pass with explicit layouts: point k of line (o, i) is read at o*in_plane + k*in_kstride + i and
 * written at o*out_plane + k*out_kstride + i (complex elements; columns i stay contiguous). Lets a pass
 * transpose while it transforms, e.g. read (o,k,i) and write (k,o,i). Out of place only unless the two
 * layouts are identical. */
int sstc_fft_axis_strided_dev(const double *d_in, double *d_out, int64_t outer, int32_t len, int64_t inner,
        int64_t in_plane, int64_t in_kstride, int64_t out_plane, int64_t out_kstride, int inverse, double scale,
        void *stream);

/* The same pass with blocked addressing on either side. The axis is cut into n equal blocks of len/n points;
 * point k of plane o, column i lives at blocks[k / (len/n)] + (o*(len/n) + k % (len/n))*pitch + i (complex
 * elements; pitch 0 = inner, i.e. each block is a dense (outer, len/n, inner) array; n = 1 with pitch 0 is
 * the plain layout). `d_*_blocks` are HOST arrays of device pointers, n <= 8. Pointing the output blocks at
 * the peers' receive buffers (P2P-mapped) turns the pass into "FFT + slab transpose over NVLink" in one
 * kernel; the input side reads an axis that an all-to-all delivered as one block per source rank. */
int sstc_fft_axis_blocked_dev(void *const *d_in_blocks, int n_in, int64_t in_pitch, void *const *d_out_blocks,
        int n_out, int64_t out_pitch, int64_t outer, int32_t len, int64_t inner, int inverse, double scale,
        void *stream);

/* Contiguous-axis pass with a line scatter. The input is (planes, plane_lines, len); the lines of a plane are
 * cut into `nblocks` equal blocks. The call transforms lines [first, first + group) of every block of every
 * plane and stores line yl of block q of plane xl at d_out_blocks[q] + (xl*(plane_lines/nblocks) + yl)*len
 * (complex elements): whole lines per destination. This is the exchange pass of the pipelined slab transform:
 * issued group by group, so the next axis can start on a group while later groups are still in flight.
 * max_ctas > 0 caps the grid (the CTAs then loop over the tiles), leaving SM capacity for a concurrent pass. */
int sstc_fft_lines_scatter_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes,
        int32_t plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas,
        void *stream);

/* The same with the store variant selectable: bulk = 1 puts every transformed line back into shared memory and sends it
 * as one bulk asynchronous copy (cp.async.bulk shared -> global) instead of per-thread 16-byte stores. */
int sstc_fft_lines_scatter_ex_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int64_t planes,
        int32_t plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale, int64_t max_ctas, int bulk,
        void *stream);

/* The mirrored line scatter, for the inverse slab transform. The input is (nblocks, block_planes, plane_lines, len):
 * block q holds the planes that belong to rank q. The call transforms every line of planes [first, first + group) of
 * every block and stores line yl of plane xl of block q at d_out_blocks[q] + (xl*dst_plane_lines + yl)*len, i.e.
 * into a destination whose planes have dst_plane_lines lines (d_out_blocks[q] already points at this rank's first
 * line inside a plane). Issued plane group by plane group so that the receiver can start its last axis on a group
 * while later groups are in flight. */
int sstc_fft_planes_scatter_dev(const double *d_in, void *const *d_out_blocks, int nblocks, int32_t block_planes,
        int32_t plane_lines, int32_t dst_plane_lines, int32_t len, int32_t first, int32_t group, int inverse, double scale,
        int64_t max_ctas, void *stream);

/* ---- slab-decomposed 3-D transform over the GPUs of one box (the reference has no distributed
 * path, FftService.java:38-106 is what it extends) ------------------------------------------------------------------
 *
 * dims (d0, d1, d2) complex128, P ranks, P | d0, P | d1, d2 a power of two in [8, 4096], d0 and d1 <= 4096 (or any length
 * <= 7168). "Natural" slabs: rank g holds x[g*d0/P : (g+1)*d0/P, :, :]; "transposed" slabs: rank g holds
 * X[:, g*d1/P : (g+1)*d1/P, :] as a dense (d0, d1/P, d2) array. forward: natural -> transposed, unscaled; inverse:
 * transposed -> natural, scaled by 1/N. The schedule (y pass locally, z pass storing whole lines into the owners'
 * buffers over NVLink group by group, flag barrier, x pass of each group beside the rest of the exchange) runs on the
 * context's own streams, joined to `stream` on both sides, and is replayed from a CUDA graph after two eager calls with
 * the same input pointer.
 *
 * sstc_slab: ONE rank's context, for the one-process-per-GPU layout. d_recv0 / d_recv1 [nranks]: every rank's two receive
 * buffers (d0*d1*d2/P complex each) as mapped in THIS process (sstc_ipc_open; entry `rank` is this rank's own
 * allocation); d_flags [nranks]: every rank's zero-initialised flag array (16 uint64). All ranks must issue the same
 * sequence of forward / inverse calls. *d_out receives this rank's result slab inside its own receive buffer: valid
 * until the next-but-one call on this context, provided its consumers are enqueued on `stream` before the next call.
 * A peer that never arrives makes a later call fail with "peer barrier timed out" instead of returning stale data. */
typedef struct sstc_slab_s *sstc_slab;
int sstc_slab_create(sstc_slab *slab, const int32_t *dims, int nranks, int rank, void *const *d_recv0, void *const *d_recv1,
        void *const *d_flags);
int sstc_slab_forward(sstc_slab slab, const double *d_in, double **d_out, void *stream);
int sstc_slab_inverse(sstc_slab slab, const double *d_in, double **d_out, void *stream);
/* For a STREAM of independent transforms: issues the local y pass of a transform now, on its own stream, so that it runs
 * beside the exchange of the transform in flight; the next sstc_slab_forward with the same d_in finishes it (exchange and
 * x passes only). At most two transforms may be begun ahead; they are finished in the order they were begun, and no
 * inverse may be issued in between. A caller that overlaps this way keeps d_in unchanged until the matching forward. */
int sstc_slab_forward_begin(sstc_slab slab, const double *d_in, void *stream);
/* Options: "groups" (pipeline groups of the exchange, default 4), "chunks" (the local pass in front of the exchange is
 * cut into this many pieces and the first group follows piece by piece; default 4 at 8 ranks, else 1), "exchange_ctas"
 * (grid cap of the exchange pass; default one CTA per SM, two at 2 ranks; 0 = uncapped), "bulk_store" (1: every
 * transformed line leaves shared memory as one cp.async.bulk copy instead of per-thread stores), "graph" (0: always
 * issue eagerly). Setting an option drops the recorded graphs. */
int sstc_slab_set_option(sstc_slab slab, const char *name, int64_t value);
int64_t sstc_slab_graph_count(sstc_slab slab);
/* Diagnostics: with option "trace" = 1 steps are issued eagerly with a timing event behind every launch; after a
 * synchronise this writes "label<TAB>ms since the step's start<NEWLINE>" records into buf (at most max bytes) and
 * returns the bytes written. tools/trace_sharded.py prints the timeline of one step on every rank. */
int64_t sstc_slab_trace(sstc_slab slab, char *buf, int64_t max);
int sstc_slab_destroy(sstc_slab slab);

/* The same transform driven from ONE process over `ndev` devices (devices == NULL: 0..ndev-1): the library allocates
 * the receive buffers and flags on every device and enables peer access (cudaDeviceEnablePeerAccess). d_in[r] / d_out[r]
 * are rank r's slabs on devices[r]; streams[r] (or NULL: the plan's own stream per device) is the stream rank r's work
 * is joined to. Calls return without synchronising; sstc_fft3d_sharded_sync waits for every device.
 * sstc_fft3d_sharded_host is the HOST tier: whole (d0, d1, d2) arrays in host memory, natural order in and out; every
 * device moves its slab over its own PCIe link. This is what CudaFftService.setHint("devices", ...) binds. */
typedef struct sstc_fft3d_sharded_s *sstc_fft3d_sharded;
int sstc_fft3d_sharded_create(sstc_fft3d_sharded *plan, const int32_t *dims, int ndev, const int32_t *devices);
int sstc_fft3d_sharded_forward(sstc_fft3d_sharded plan, const double *const *d_in, double **d_out, void *const *streams);
int sstc_fft3d_sharded_inverse(sstc_fft3d_sharded plan, const double *const *d_in, double **d_out, void *const *streams);
int sstc_fft3d_sharded_sync(sstc_fft3d_sharded plan);
int sstc_fft3d_sharded_set_option(sstc_fft3d_sharded plan, const char *name, int64_t value);
int sstc_fft3d_sharded_host(sstc_fft3d_sharded plan, int inverse, const double *in, double *out);
int sstc_fft3d_sharded_destroy(sstc_fft3d_sharded plan);

/* Cross-GPU barrier kernel (one process per GPU): d_flag_arrays[r] is rank r's array of 8 uint64 epoch slots,
 * mapped by every rank (sstc_ipc_*) and zero-initialised (at least 9 slots: slot 8 is the rank's own call
 * counter); `epoch` must increase with every call, or be 0 to let the kernel count calls itself (for use inside
 * CUDA graphs). Stream-ordered:
 * peer stores issued by earlier kernels on `stream` are complete before the flag is published, and kernels
 * after it start only when every rank has published `epoch`. */
int sstc_peer_barrier_dev(void *const *d_flag_arrays, int nranks, int rank, uint64_t epoch, void *stream);

/* Diagnostics (tools/exp_nvlink.py): copies `bytes` from src to dst (either may be a peer's memory) with `ctas` CTAs.
 * mode 0: per-thread 16-byte loads and stores (the exchange pass's pattern); mode 1: bulk asynchronous copies through
 * shared memory in pieces of `piece` bytes (cp.async.bulk in, cp.async.bulk out). Not used by any transform. */
int sstc_exp_peer_copy(void *dst, const void *src, int64_t bytes, int mode, int piece, int ctas, void *stream);

/* Diagnostics knobs: "exchange_smem_kb" = dynamic shared memory the pipelined exchange kernel requests (beyond its need),
 * which keeps other kernels' CTAs off its SMs (tools/exp_overlap.py); "tma_strided" = 1 / 2 / 3: the 512- and 4096-point
 * strided passes move their tiles with TMA (cp.async.bulk.tensor) with that many tile buffers per CTA instead of per-thread
 * loads and stores (tools/exp_tma.py; 0 = off, the default). */
int sstc_exp_set(const char *name, int64_t value);

/* Cross-process device-memory sharing for the one-process-per-GPU layout (cudaIpc*): export a 64-byte
 * handle for a buffer from sstc_malloc, open a peer's handle, close it. */
#define SSTC_IPC_HANDLE_BYTES 64
int sstc_ipc_export(void *dptr, unsigned char handle[SSTC_IPC_HANDLE_BYTES]);
int sstc_ipc_open(const unsigned char handle[SSTC_IPC_HANDLE_BYTES], void **dptr);
int sstc_ipc_close(void *dptr);

/* ---- ArrayKernel ------------------------------------------------------------------------------------
 *
 * One HOST-tier entry point per native method of org.shared.array.jni.NativeArrayKernel
 * (src/org/shared/array/jni/NativeArrayKernel.java:56-154, JNI exports native/src/shared/Bindings.cpp:26-142),
 * with the same argument order; a Java array becomes (pointer, length). `type` is the ArrayKernel opcode
 * (src/org/shared/array/kernel/ArrayKernel.java:45-278). Validation and error messages follow the reference's
 * C++ ("Invalid arguments", "Invalid dimensions and/or strides", "Invalid array lengths", "Operation type not
 * recognized", "Duplicate operating dimensions are not allowed", "Operating dimensions must have singleton or
 * zero length", "Invalid dimension", "Invalid index", "Invalid mapping parameters", "Invalid array types").
 * dims / strides / bounds / slices / opDims are small HOST arrays in BOTH tiers. Every `_dev` twin takes device
 * pointers for the value arrays plus a stream, and returns without synchronising (scratch is per host thread:
 * issue one thread's calls on one stream). */

/* element kinds for eOp (ElementOps.cpp:285-321 dispatches on the runtime array type and the complex flag) */
#define SSTC_ELEM_REAL 0
#define SSTC_ELEM_COMPLEX 1
#define SSTC_ELEM_INT 2
/* conversion families for convert (ElementOps.cpp:323-407) */
#define SSTC_CONV_R_TO_C 0
#define SSTC_CONV_C_TO_R 1
#define SSTC_CONV_I_TO_R 2

/* randomize / derandomize (Bindings.cpp:26-32; JavaArrayKernel.java:57-64): reseeds RND and SHUFFLE. */
int sstc_srand(uint64_t seed);

/* ruOp / cuOp / iuOp: in-place unary maps (Bindings.cpp:86-99). n = array length (doubles / ints). */
int sstc_ruop(int type, double a, double *v, int64_t n);
int sstc_cuop(int type, double re, double im, double *v, int64_t n);
int sstc_iuop(int type, int32_t a, int32_t *v, int64_t n);
int sstc_ruop_dev(int type, double a, double *v, int64_t n, void *stream);
int sstc_cuop_dev(int type, double re, double im, double *v, int64_t n, void *stream);
int sstc_iuop_dev(int type, int32_t a, int32_t *v, int64_t n, void *stream);

/* eOp: dst = lhs (op) rhs, n = length of each array in its storage type; dst may alias lhs (Bindings.cpp:101-104). */
int sstc_eop(int type, int elem, const void *lhs, const void *rhs, void *dst, int64_t n);
int sstc_eop_dev(int type, int elem, const void *lhs, const void *rhs, void *dst, int64_t n, void *stream);

/* convert (Bindings.cpp:106-109). */
int sstc_convert(int type, int conv, const void *src, int64_t nsrc, double *dst, int64_t ndst);
int sstc_convert_dev(int type, int conv, const void *src, int64_t nsrc, double *dst, int64_t ndst, void *stream);

/* raOp -> one double, caOp -> {re, im} (Bindings.cpp:76-84). The _dev twins write to DEVICE memory. */
int sstc_raop(int type, const double *v, int64_t n, double *out);
int sstc_caop(int type, const double *v, int64_t n, double out[2]);
int sstc_raop_dev(int type, const double *v, int64_t n, double *d_out, void *stream);
int sstc_caop_dev(int type, const double *v, int64_t n, double *d_out2, void *stream);

/* rrOp: reduce over opDims (sorted in place, as the reference does); dst keeps size-1 dims (Bindings.cpp:54-62). */
int sstc_rrop(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, int32_t *opDims, int nop);
int sstc_rrop_dev(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, int32_t *opDims, int nop, void *stream);

/* riOp: positions along `dim` (or dim = -1: whole array); RI_SORT also sorts src in place (Bindings.cpp:64-70). */
int sstc_riop(int type, double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, int32_t *dst,
        int64_t ndst, int dim);
int sstc_riop_dev(int type, double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, int32_t *dst,
        int64_t ndst, int dim, void *stream);

/* rdOp: inclusive prefix scan along each of opDims in ascending order (Bindings.cpp:72-74). */
int sstc_rdop(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, double *dst,
        int64_t ndst, const int32_t *opDims, int nop);
int sstc_rdop_dev(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, double *dst,
        int64_t ndst, const int32_t *opDims, int nop, void *stream);

/* map / slice (Bindings.cpp:34-52): elem_bytes = 8 (double[]) or 4 (int[]); Object[] arrays stay on the JVM side. */
int sstc_map(int elem_bytes, const int32_t *bounds, int nbounds, const void *src, int64_t nsrc, const int32_t *sD,
        const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd);
int sstc_map_dev(int elem_bytes, const int32_t *bounds, int nbounds, const void *src, int64_t nsrc, const int32_t *sD,
        const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, void *stream);
int sstc_slice(int elem_bytes, const int32_t *slices, int nslices3, const void *src, int64_t nsrc, const int32_t *sD,
        const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd);
int sstc_slice_dev(int elem_bytes, const int32_t *slices, int nslices3, const void *src, int64_t nsrc,
        const int32_t *sD, const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd,
        void *stream);

/* The structural callers of map / slice as single device-tier calls on dense arrays (near_order = 0: row-major FAR,
 * 1: column-major NEAR; Array.java:240-287). In the reference they live on the JVM side and build an int[] specification
 * per call: ProtoArray.shift (ProtoArray.java:313-337: element i of dimension d moves to (i + shifts[d]) mod dims[d]),
 * AbstractComplexArray.fftShift / ifftShift (AbstractComplexArray.java:67-97: dims includes the trailing 2; every other
 * dimension is shifted by +-dims[d]/2), ProtoArray.transpose (ProtoArray.java:246-283: source dimension d becomes
 * destination dimension permutation[d]; "Invalid permutation"), ConvolutionCache.pad (ConvolutionCache.java:149-239:
 * mirror padding by margins[d] on both sides of dimension d, row-major; dst has dims[d] + 2 margins[d]). src != dst. */
int sstc_shift_dev(int elem_bytes, const void *src, void *dst, const int32_t *dims, int nd, int near_order, const int32_t *shifts,
        void *stream);
int sstc_fftshift_dev(const double *src, double *dst, const int32_t *dims, int nd, int inverse, void *stream);
int sstc_transpose_dev(int elem_bytes, const void *src, void *dst, const int32_t *dims, int nd, int near_order,
        const int32_t *permutation, void *stream);
int sstc_pad_dev(const double *src, double *dst, const int32_t *dims, int nd, const int32_t *margins, void *stream);

/* mul / diag (Bindings.cpp:111-119): row-major, inner dimension inferred from the lengths. */
int sstc_mul(const double *lhs, int64_t nl, const double *rhs, int64_t nr, int lr, int rc, double *dst, int64_t ndst,
        int complex);
int sstc_mul_dev(const double *lhs, int64_t nl, const double *rhs, int64_t nr, int lr, int rc, double *dst, int64_t ndst,
        int complex, void *stream);
int sstc_diag(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size, int complex);
int sstc_diag_dev(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size, int complex, void *stream);

/* invert (Bindings.cpp:139-142 -> LinearAlgebraOpsLu.cpp:26-76): dst = src^-1 for a row-major size x size matrix, by
 * LU with partial pivoting (JAMA's pivot rule). "Matrix is singular" when U has a zero on its diagonal. dst is
 * overwritten (the reference assumes it arrives zero-filled). The device twin synchronises `stream` once. */
int sstc_invert(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size);
int sstc_invert_dev(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size, void *stream);

/* svd (Bindings.cpp:127-132 -> LinearAlgebraOpsSvd.cpp:26-72): thin SVD of an n_rows x n_cols matrix, n_rows >= n_cols,
 * given row-major (strides n_cols, 1) or as the transpose of a row-major matrix (strides 1, n_rows). Outputs row-major
 * U (n_rows x n_cols), S (n_cols, descending), V (n_cols x n_cols) with U diag(S) V^T = A. Computed by one-sided Jacobi
 * rotations, so U and V agree with the reference's JAMA port up to the sign of each singular-vector pair. */
int sstc_svd(const double *src, int64_t nsrc, int src_stride_row, int src_stride_col, double *u, int64_t nu, double *sv,
        int64_t ns, double *v, int64_t nv, int n_rows, int n_cols);
int sstc_svd_dev(const double *src, int64_t nsrc, int src_stride_row, int src_stride_col, double *u, int64_t nu, double *sv,
        int64_t ns, double *v, int64_t nv, int n_rows, int n_cols, void *stream);

/* eigs (Bindings.cpp:134-137 -> LinearAlgebraOpsEigs.cpp:26-615; ArrayKernel.java:574): eigenvectors `vec` (row-major
 * size x size, one vector per column; a complex pair occupies two adjacent columns: real part, imaginary part) and
 * eigenvalues `val` (2 * size doubles, interleaved re / im) of a general real row-major size x size matrix: JAMA's
 * Hessenberg reduction + double-shift QR + back-substitution, every element updated in the reference's operation
 * order, so the output is bit-identical to the reference's. src is not modified. Fails with "Invalid arguments" on a
 * length mismatch and with "Eigenvalue iteration did not converge" where the reference's uncapped hqr2 loop would
 * never return (an all-zero matrix of size >= 3, NaN input). The device twin synchronises `stream` once. */
int sstc_eigs(const double *src, int64_t nsrc, double *vec, int64_t nvec, double *val, int64_t nval, int size);
int sstc_eigs_dev(const double *src, int64_t nsrc, double *vec, int64_t nvec, double *val, int64_t nval, int size,
        void *stream);

/* find (Bindings.cpp:160-163 -> IndexOps.cpp:26-130): `logical` holds one index per dimension with exactly one -1
 * marking the active dimension; along that line of the int array (as riOp writes it: positions, then -1 padding)
 * *count = number of non-negative entries, and out[0 .. *count) = the first *count entries of the line. `out` must
 * hold a whole line (sD[active] ints; host tier: nout says how many it holds). The device twin takes device
 * `src` / `out`, writes *count on the host and synchronises `stream`. */
int sstc_find(const int32_t *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, const int32_t *logical,
        int32_t *out, int64_t nout, int32_t *count);
int sstc_find_dev(const int32_t *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, const int32_t *logical,
        int32_t *out, int32_t *count, void *stream);

/* ---- sparse arrays (Bindings.cpp:165-182 -> SparseOpsInsert.cpp:26-168, SparseOpsSlice.cpp:26-410, SparseOps.cpp:26-260;
 * ArrayKernel.java:628-674) --------------------------------------------------------------------------------
 *
 * Both calls return the four arrays of org.shared.array.sparse.SparseArrayState: `values` (count elements of
 * elem_bytes), `indices` (count ascending physical indices), `indirection_offsets` (sum(dims) + nd entries: per
 * dimension the bucket starts of each logical index, closed by count) and `indirections` (nd * count: per dimension the
 * element positions ordered by logical index, ties in ascending position). The pointers refer to library-owned host
 * memory that stays valid until the calling thread's next sparse call; the JNI glue copies them into new Java arrays.
 * elem_bytes: 8 (double[]) or 4 (int[]); the glue serves Object[] by passing positions as 4-byte values and gathering
 * the references afterwards. Sorting, merging, metadata and value movement run on the device. Bit-exact. */
typedef struct sstc_sparse_state {
    int32_t count, n_offsets, n_indirections, elem_bytes;
    const void *values;
    const int32_t *indices, *indirection_offsets, *indirections;
} sstc_sparse_state;

/* insertSparse: merges new (value, physical index) pairs into a sparse array; on equal indices the new value wins.
 * oldI is ascending without duplicates; newI is in any order and comes back SORTED, as with both reference providers
 * (SparseOpsInsert.cpp:118-126, SparseOps.java:80-84). oldDo has nd + 1 entries with oldDo[d + 1] - oldDo[d] - 1 ==
 * oldD[d]. Errors: "Invalid arguments", "Invalid dimensions and/or strides", "Duplicate values are not allowed",
 * "Invalid physical index", "Invalid array types". */
int sstc_insert_sparse(int elem_bytes, const void *oldV, int64_t nold, const int32_t *oldD, const int32_t *oldS,
        const int32_t *oldDo, int nd, const int32_t *oldI, const void *newV, int64_t nnew, int32_t *newI,
        sstc_sparse_state *out);

/* sliceSparse: `slices` holds nslices3 / 3 triples (source index, destination index, dimension). The destination
 * cells whose logical index is a sliced destination index in EVERY dimension are cleared; every source element whose
 * logical index is a sliced source index in every dimension is written to the cross product of the destination
 * indices its coordinates map to (the lowest source position wins where two collide); the rest of the destination is
 * kept. srcIo / dstIo (sum(dims) + nd entries) are range-checked for the sliced indices like the reference does;
 * srcIi / dstIi must be present but the selection is recomputed from the physical indices, so they are not read.
 * Errors as above plus "Invalid dimension", "Invalid index". */
int sstc_slice_sparse(int elem_bytes, const int32_t *slices, int nslices3, const void *srcV, int64_t nsrc,
        const int32_t *srcD, const int32_t *srcS, const int32_t *srcDo, const int32_t *srcI, const int32_t *srcIo,
        const int32_t *srcIi, const void *dstV, int64_t ndst, const int32_t *dstD, const int32_t *dstS,
        const int32_t *dstDo, const int32_t *dstI, const int32_t *dstIo, const int32_t *dstIi, int nd,
        sstc_sparse_state *out);

/* ---- ImageKernel ------------------------------------------------------------------------------------
 *
 * The reference library's second native service (org.shared.image.kernel.ImageKernel,
 * src/org/shared/image/kernel/ImageKernel.java:56-80; native methods org.shared.image.jni.NativeImageKernel,
 * JNI exports native/src/shared/Bindings.cpp:144-158, bodies native/src/shared/NativeImageKernel.cpp:32-247;
 * registered by JNI_OnLoad next to the ArrayKernel, native/src/shared_common/Library.cpp:64-66).
 * dst has every source dimension + 1 (histogram: plus a trailing bins dimension of dD[nd] entries); the source
 * lands at offset +1 along every dimension (histogram: in the bin mem[] names), the rest of dst is left as
 * passed, and an inclusive running sum is taken along dimensions 0..nd-1 in turn. Errors: "Invalid arguments",
 * "Invalid dimensions and/or strides", "Dimension mismatch", "Invalid membership index". */
int sstc_integral_image(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd);
int sstc_integral_image_dev(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, void *stream);
/* sD / sS have nd entries, dD / dS nd + 1; mem[] (int32 class memberships) is addressed like src. The device
 * twin synchronises `stream` once (the membership check must finish before dst is touched). */
int sstc_integral_histogram(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd,
        const int32_t *mem, int64_t nmem, double *dst, int64_t ndst, const int32_t *dD, const int32_t *dS);
int sstc_integral_histogram_dev(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd,
        const int32_t *mem, int64_t nmem, double *dst, int64_t ndst, const int32_t *dD, const int32_t *dS,
        void *stream);

#ifdef __cplusplus
}
#endif

#endif /* SST_CUDA_H */

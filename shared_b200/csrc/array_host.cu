// array_host.cu -- HOST tier of the ArrayKernel C ABI: the entry points the JNI glue binds one-to-one to the
// methods of org.shared.array.jni.NativeArrayKernel (native/src/shared/Bindings.cpp:26-142). Pointers are the
// pinned Java arrays; each call stages them through HBM (H2D, device-tier kernels, D2H of whatever the Java
// contract says is mutated) on the legacy default stream and returns when the host arrays are up to date.
// Where the reference pins with JNI_ABORT (inputs, no copy-back) nothing is copied back here either.
#include <algorithm>
#include <string>

#include "array_common.cuh"

namespace sstc {

// Staging buffers are per host thread, like the reference's per-call MallocHandler scratch.
static thread_local Scratch t_stage[3];

struct Staged {
    void *dev = nullptr;
    void *host = nullptr;
    int64_t bytes = 0;
    Staged(int slot, const void *h, int64_t nbytes, bool upload) : host(const_cast<void *>(h)), bytes(nbytes) {
        if (nbytes > 0) {
            dev = t_stage[slot].get(nbytes);
            if (upload) {
                SSTC_CUDA(cudaMemcpyAsync(dev, h, (size_t) nbytes, cudaMemcpyHostToDevice, 0));
            }
        }
    }
    void download() {
        if (bytes > 0) {
            SSTC_CUDA(cudaMemcpyAsync(host, dev, (size_t) bytes, cudaMemcpyDeviceToHost, 0));
        }
    }
};

static void finish() { SSTC_CUDA(cudaStreamSynchronize(0)); }

static void rethrow(int rc) {
    if (rc != 0) {
        throw std::runtime_error(std::string(sstc_last_error()));
    }
}

}  // namespace sstc

using namespace sstc;

extern "C" {

int sstc_ruop(int type, double a, double *v, int64_t n) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !v)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, v, n * 8, true);
        rethrow(sstc_ruop_dev(type, a, static_cast<double *>(s0.dev), n, nullptr));
        s0.download();
        finish();
    });
}

int sstc_cuop(int type, double re, double im, double *v, int64_t n) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !v)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, v, n * 8, true);
        rethrow(sstc_cuop_dev(type, re, im, static_cast<double *>(s0.dev), n, nullptr));
        s0.download();
        finish();
    });
}

int sstc_iuop(int type, int32_t a, int32_t *v, int64_t n) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !v)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, v, n * 4, true);
        rethrow(sstc_iuop_dev(type, a, static_cast<int32_t *>(s0.dev), n, nullptr));
        s0.download();
        finish();
    });
}

int sstc_eop(int type, int elem, const void *lhs, const void *rhs, void *dst, int64_t n) {
    return guarded([&] {
        if (elem < SSTC_ELEM_REAL || elem > SSTC_ELEM_INT) {
            throw std::runtime_error("Array type not recognized");
        }
        if (n < 0 || (n > 0 && (!lhs || !rhs || !dst))) {
            throw std::runtime_error("Invalid arguments");
        }
        const int64_t eb = elem == SSTC_ELEM_INT ? 4 : 8;
        // dst may alias lhs (the l* operations, AbstractRealArray.java:511-520): three device buffers, one download
        Staged l(0, lhs, n * eb, true), r(1, rhs, n * eb, true), d(2, dst, n * eb, false);
        rethrow(sstc_eop_dev(type, elem, l.dev, r.dev, d.dev, n, nullptr));
        d.download();
        finish();
    });
}

int sstc_convert(int type, int conv, const void *src, int64_t nsrc, double *dst, int64_t ndst) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * (conv == SSTC_CONV_I_TO_R ? 4 : 8), true), d(1, dst, ndst * 8, false);
        rethrow(sstc_convert_dev(type, conv, s0.dev, nsrc, static_cast<double *>(d.dev), ndst, nullptr));
        d.download();
        finish();
    });
}

int sstc_raop(int type, const double *v, int64_t n, double *out) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !v) || !out) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, v, n * 8, true), d(1, out, 8, false);
        rethrow(sstc_raop_dev(type, static_cast<const double *>(s0.dev), n, static_cast<double *>(d.dev), nullptr));
        d.download();
        finish();
    });
}

int sstc_caop(int type, const double *v, int64_t n, double out[2]) {
    return guarded([&] {
        if (n < 0 || (n > 0 && !v) || !out) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, v, n * 8, true), d(1, out, 16, false);
        rethrow(sstc_caop_dev(type, static_cast<const double *>(s0.dev), n, static_cast<double *>(d.dev), nullptr));
        d.download();
        finish();
    });
}

int sstc_rrop(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd, int32_t *opDims, int nop) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        // dst is uploaded too: positions the (possibly strided) destination layout does not cover keep their values
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 8, true);
        rethrow(sstc_rrop_dev(type, static_cast<const double *>(s0.dev), nsrc, sD, sS, static_cast<double *>(d.dev), ndst,
                dD, dS, nd, opDims, nop, nullptr));
        d.download();
        finish();
    });
}

int sstc_riop(int type, double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, int32_t *dst,
        int64_t ndst, int dim) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 4, true);
        rethrow(sstc_riop_dev(type, static_cast<double *>(s0.dev), nsrc, sD, sS, nd, static_cast<int32_t *>(d.dev), ndst,
                dim, nullptr));
        d.download();
        if (type == 5) {
            s0.download();  // RI_SORT sorts the source in place (DimensionOpsIndex.cpp:80 pins it READ_WRITE)
        }
        finish();
    });
}

int sstc_rdop(int type, const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, double *dst,
        int64_t ndst, const int32_t *opDims, int nop) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 8, false);
        rethrow(sstc_rdop_dev(type, static_cast<const double *>(s0.dev), nsrc, sD, sS, nd, static_cast<double *>(d.dev),
                ndst, opDims, nop, nullptr));
        d.download();
        finish();
    });
}

int sstc_map(int elem_bytes, const int32_t *bounds, int nbounds, const void *src, int64_t nsrc, const int32_t *sD,
        const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd) {
    return guarded([&] {
        if (elem_bytes != 4 && elem_bytes != 8) {
            throw std::runtime_error("Invalid array types");
        }
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * elem_bytes, true), d(1, dst, ndst * elem_bytes, true);
        rethrow(sstc_map_dev(elem_bytes, bounds, nbounds, s0.dev, nsrc, sD, sS, d.dev, ndst, dD, dS, nd, nullptr));
        d.download();
        finish();
    });
}

int sstc_slice(int elem_bytes, const int32_t *slices, int nslices3, const void *src, int64_t nsrc, const int32_t *sD,
        const int32_t *sS, void *dst, int64_t ndst, const int32_t *dD, const int32_t *dS, int nd) {
    return guarded([&] {
        if (elem_bytes != 4 && elem_bytes != 8) {
            throw std::runtime_error("Invalid array types");
        }
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * elem_bytes, true), d(1, dst, ndst * elem_bytes, true);
        rethrow(sstc_slice_dev(elem_bytes, slices, nslices3, s0.dev, nsrc, sD, sS, d.dev, ndst, dD, dS, nd, nullptr));
        d.download();
        finish();
    });
}

int sstc_mul(const double *lhs, int64_t nl, const double *rhs, int64_t nr, int lr, int rc, double *dst, int64_t ndst,
        int complex) {
    return guarded([&] {
        if (nl < 0 || nr < 0 || ndst < 0 || (nl > 0 && !lhs) || (nr > 0 && !rhs) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged l(0, lhs, nl * 8, true), r(1, rhs, nr * 8, true), d(2, dst, ndst * 8, false);
        rethrow(sstc_mul_dev(static_cast<const double *>(l.dev), nl, static_cast<const double *>(r.dev), nr, lr, rc,
                static_cast<double *>(d.dev), ndst, complex, nullptr));
        d.download();
        finish();
    });
}

int sstc_diag(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size, int complex) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 8, false);
        rethrow(sstc_diag_dev(static_cast<const double *>(s0.dev), nsrc, static_cast<double *>(d.dev), ndst, size, complex,
                nullptr));
        d.download();
        finish();
    });
}

// ---- ImageKernel (org.shared.image.jni.NativeImageKernel, Bindings.cpp:144-158) -----------------------------

int sstc_integral_image(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, double *dst,
        int64_t ndst, const int32_t *dD, const int32_t *dS, int nd) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        // dst is uploaded too: its border takes part in the sums as it is (NativeImageKernel.cpp:96-118)
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 8, true);
        rethrow(sstc_integral_image_dev(static_cast<const double *>(s0.dev), nsrc, sD, sS, static_cast<double *>(d.dev),
                ndst, dD, dS, nd, nullptr));
        d.download();
        finish();
    });
}

int sstc_integral_histogram(const double *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd,
        const int32_t *mem, int64_t nmem, double *dst, int64_t ndst, const int32_t *dD, const int32_t *dS) {
    return guarded([&] {
        if (nsrc < 0 || nmem < 0 || ndst < 0 || (nsrc > 0 && !src) || (nmem > 0 && !mem) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 8, true), m(2, mem, nmem * 4, true), d(1, dst, ndst * 8, true);
        rethrow(sstc_integral_histogram_dev(static_cast<const double *>(s0.dev), nsrc, sD, sS, nd,
                static_cast<const int32_t *>(m.dev), nmem, static_cast<double *>(d.dev), ndst, dD, dS, nullptr));
        d.download();
        finish();
    });
}

// ---- invert / find (Bindings.cpp:139-142, 160-163) --------------------------------------------------------------

int sstc_invert(const double *src, int64_t nsrc, double *dst, int64_t ndst, int size) {
    return guarded([&] {
        if (nsrc < 0 || ndst < 0 || (nsrc > 0 && !src) || (ndst > 0 && !dst)) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 8, true), d(1, dst, ndst * 8, false);
        rethrow(sstc_invert_dev(static_cast<const double *>(s0.dev), nsrc, static_cast<double *>(d.dev), ndst, size,
                nullptr));
        d.download();
        finish();
    });
}

int sstc_svd(const double *src, int64_t nsrc, int src_stride_row, int src_stride_col, double *u, int64_t nu, double *sv,
        int64_t ns, double *v, int64_t nv, int n_rows, int n_cols) {
    return guarded([&] {
        if (nsrc < 0 || nu < 0 || ns < 0 || nv < 0 || (nsrc > 0 && !src) || (nu > 0 && !u) || (ns > 0 && !sv) ||
                (nv > 0 && !v)) {
            throw std::runtime_error("Invalid arguments");
        }
        static thread_local Scratch extra;  // a fourth staging buffer: src, U, S, V
        Staged s0(0, src, nsrc * 8, true), du(1, u, nu * 8, false), dv(2, v, nv * 8, false);
        double *ds = static_cast<double *>(extra.get(std::max<int64_t>(ns, 1) * 8));
        rethrow(sstc_svd_dev(static_cast<const double *>(s0.dev), nsrc, src_stride_row, src_stride_col,
                static_cast<double *>(du.dev), nu, ds, ns, static_cast<double *>(dv.dev), nv, n_rows, n_cols, nullptr));
        du.download();
        dv.download();
        if (ns > 0) {
            SSTC_CUDA(cudaMemcpyAsync(sv, ds, (size_t) ns * 8, cudaMemcpyDeviceToHost, 0));
        }
        finish();
    });
}

int sstc_find(const int32_t *src, int64_t nsrc, const int32_t *sD, const int32_t *sS, int nd, const int32_t *logical,
        int32_t *out, int64_t nout, int32_t *count) {
    return guarded([&] {
        if (nsrc < 0 || nout < 0 || (nsrc > 0 && !src) || (nout > 0 && !out) || !count) {
            throw std::runtime_error("Invalid arguments");
        }
        Staged s0(0, src, nsrc * 4, true), d(1, out, nout * 4, false);
        for (int i = 0; i < nd && sD && logical; i++) {
            if (logical[i] == -1 && nout < sD[i]) {
                throw std::runtime_error("Invalid arguments");  // `out` must hold a whole line
            }
        }
        rethrow(sstc_find_dev(static_cast<const int32_t *>(s0.dev), nsrc, sD, sS, nd, logical,
                static_cast<int32_t *>(d.dev), count, nullptr));
        if (*count > 0) {
            SSTC_CUDA(cudaMemcpyAsync(out, d.dev, (size_t) *count * 4, cudaMemcpyDeviceToHost, 0));
        }
        finish();
    });
}

}  // extern "C"

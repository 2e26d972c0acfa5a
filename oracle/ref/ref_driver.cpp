/*
 * TEST INFRASTRUCTURE ONLY -- flat C entry points over the reference's own native ArrayKernel.
 *
 * oracle/Makefile compiles this file together with the UNMODIFIED sources
 * /root/reference/native/src/shared/{Common,ElementOps,DimensionOps*,MappingOps*,MatrixOps,NativeArrayKernel,
 * IndexOps,NativeImageKernel}.cpp (read where they lie; never copied) against the shim oracle/ref/jni.h into
 * oracle/_ref/libsst_ref.so. The parity tests call these wrappers; each one builds the {kind,data,len} array
 * records the shim uses for Java arrays and forwards to the function the JNI export table binds
 * (native/src/shared/Bindings.cpp:26-142). A Java exception shows up as return code 1 + sstref_last_error().
 */
#include <Bindings.hpp>

#include <stdlib.h>
#include <string.h>

#include <string>

static std::string g_error;
static bool g_init = false;

static JNIEnv *env_get() {
    static JNIEnv env;
    if (!g_init) {
        NativeArrayKernel::init(&env); /* caches the [D / [I class handles (NativeArrayKernel.cpp:37-57) */
        g_init = true;
    }
    env.pending = false;
    env.message.clear();
    return &env;
}

static int env_done(JNIEnv *env) {
    if (env->pending) {
        g_error = env->message;
        return 1;
    }
    g_error.clear();
    return 0;
}

static _jobject arr(int kind, const void *data, int len) {
    _jobject o;
    o.kind = kind;
    o.data = const_cast<void *>(data);
    o.len = len;
    o.owned = 0;
    return o;
}

#define DARR(p, n) arr(SHIM_DOUBLE_ARRAY, p, n)
#define IARR(p, n) arr(SHIM_INT_ARRAY, p, n)

extern "C" {

const char *sstref_last_error() { return g_error.c_str(); }

void sstref_srand(unsigned seed) { srand(seed); }

int sstref_ruop(int type, double a, double *v, int n) {
    JNIEnv *env = env_get();
    _jobject s = DARR(v, n);
    ElementOps::ruOp(env, 0, type, a, &s);
    return env_done(env);
}

int sstref_cuop(int type, double re, double im, double *v, int n) {
    JNIEnv *env = env_get();
    _jobject s = DARR(v, n);
    ElementOps::cuOp(env, 0, type, re, im, &s);
    return env_done(env);
}

int sstref_iuop(int type, int a, int *v, int n) {
    JNIEnv *env = env_get();
    _jobject s = IARR(v, n);
    ElementOps::iuOp(env, 0, type, a, &s);
    return env_done(env);
}

/* kind: SHIM_DOUBLE_ARRAY (1) or SHIM_INT_ARRAY (2) for all three operands */
int sstref_eop(int type, int kind, const void *lhs, const void *rhs, void *dst, int n, int complex) {
    JNIEnv *env = env_get();
    _jobject l = arr(kind, lhs, n), r = arr(kind, rhs, n), d = arr(kind, dst, n);
    ElementOps::eOp(env, 0, type, &l, &r, &d, complex ? JNI_TRUE : JNI_FALSE);
    return env_done(env);
}

int sstref_convert(int type, const void *src, int src_kind, int nsrc, int src_complex, double *dst, int ndst,
        int dst_complex) {
    JNIEnv *env = env_get();
    _jobject s = arr(src_kind, src, nsrc), d = DARR(dst, ndst);
    ElementOps::convert(env, 0, type, &s, src_complex ? JNI_TRUE : JNI_FALSE, &d, dst_complex ? JNI_TRUE : JNI_FALSE);
    return env_done(env);
}

int sstref_raop(int type, const double *v, int n, double *out) {
    JNIEnv *env = env_get();
    _jobject s = DARR(v, n);
    *out = ElementOps::raOp(env, 0, type, &s);
    return env_done(env);
}

int sstref_caop(int type, const double *v, int n, double out[2]) {
    JNIEnv *env = env_get();
    _jobject s = DARR(v, n);
    jdoubleArray res = ElementOps::caOp(env, 0, type, &s);
    if (res && res->len == 2) {
        out[0] = ((double *) res->data)[0];
        out[1] = ((double *) res->data)[1];
        free(res->data);
        free(res);
    }
    return env_done(env);
}

int sstref_rrop(int type, const double *src, int nsrc, const int *sD, const int *sS, int nd, double *dst, int ndst,
        const int *dD, const int *dS, int *opDims, int nop) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), d = DARR(dst, ndst), dd = IARR(dD, nd),
             ds = IARR(dS, nd), od = IARR(opDims, nop);
    DimensionOps::rrOp(env, 0, type, &s, &sd, &ss, &d, &dd, &ds, &od);
    return env_done(env);
}

int sstref_riop(int type, double *src, int nsrc, const int *sD, const int *sS, int nd, int *dst, int ndst, int dim) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), d = IARR(dst, ndst);
    DimensionOps::riOp(env, 0, type, &s, &sd, &ss, &d, dim);
    return env_done(env);
}

int sstref_rdop(int type, const double *src, int nsrc, const int *sD, const int *sS, int nd, double *dst, int ndst,
        int *opDims, int nop) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), d = DARR(dst, ndst), od = IARR(opDims, nop);
    DimensionOps::rdOp(env, 0, type, &s, &sd, &ss, &d, &od);
    return env_done(env);
}

int sstref_map(const int *bounds, int nb, const void *src, int kind, int nsrc, const int *sD, const int *sS, int nds,
        void *dst, int ndst, const int *dD, const int *dS, int ndd) {
    JNIEnv *env = env_get();
    _jobject b = IARR(bounds, nb), s = arr(kind, src, nsrc), sd = IARR(sD, nds), ss = IARR(sS, nds),
             d = arr(kind, dst, ndst), dd = IARR(dD, ndd), ds = IARR(dS, ndd);
    MappingOps::map(env, 0, &b, &s, &sd, &ss, &d, &dd, &ds);
    return env_done(env);
}

int sstref_slice(const int *slices, int nsl, const void *src, int kind, int nsrc, const int *sD, const int *sS,
        int nds, void *dst, int ndst, const int *dD, const int *dS, int ndd) {
    JNIEnv *env = env_get();
    _jobject b = IARR(slices, nsl), s = arr(kind, src, nsrc), sd = IARR(sD, nds), ss = IARR(sS, nds),
             d = arr(kind, dst, ndst), dd = IARR(dD, ndd), ds = IARR(dS, ndd);
    MappingOps::slice(env, 0, &b, &s, &sd, &ss, &d, &dd, &ds);
    return env_done(env);
}

int sstref_mul(const double *lhs, int nl, const double *rhs, int nr, int lr, int rc, double *dst, int ndst,
        int complex) {
    JNIEnv *env = env_get();
    _jobject l = DARR(lhs, nl), r = DARR(rhs, nr), d = DARR(dst, ndst);
    MatrixOps::mul(env, 0, &l, &r, lr, rc, &d, complex ? JNI_TRUE : JNI_FALSE);
    return env_done(env);
}

int sstref_diag(const double *src, int nsrc, double *dst, int ndst, int size, int complex) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), d = DARR(dst, ndst);
    MatrixOps::diag(env, 0, &s, &d, size, complex ? JNI_TRUE : JNI_FALSE);
    return env_done(env);
}

int sstref_invert(const double *src, int nsrc, double *dst, int ndst, int size) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), d = DARR(dst, ndst);
    LinearAlgebraOps::invert(env, 0, &s, &d, size);
    return env_done(env);
}

int sstref_svd(const double *src, int nsrc, int rs, int cs, double *u, int nu, double *sv, int ns, double *v, int nv, int nrows,
        int ncols) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), uu = DARR(u, nu), ss = DARR(sv, ns), vv = DARR(v, nv);
    LinearAlgebraOps::svd(env, 0, &s, rs, cs, &uu, &ss, &vv, nrows, ncols);
    return env_done(env);
}

int sstref_eigs(const double *src, int nsrc, double *vec, int nvec, double *val, int nval, int size) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), ve = DARR(vec, nvec), va = DARR(val, nval);
    LinearAlgebraOps::eigs(env, 0, &s, &ve, &va, size);
    return env_done(env);
}

/* returns the length of the result (copied into out, capacity cap), or -1 on a Java exception */
int sstref_find(const int *src, int nsrc, const int *sD, const int *sS, int nd, const int *logical, int *out, int cap) {
    JNIEnv *env = env_get();
    _jobject s = IARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), lg = IARR(logical, nd);
    jintArray res = IndexOps::find(env, 0, &s, &sd, &ss, &lg);
    int n = -1;
    if (!env->pending && res) {
        n = res->len;
        for (int i = 0; i < n && i < cap; i++) {
            out[i] = ((int *) res->data)[i];
        }
    }
    if (res) {
        free(res->data);
        free(res);
    }
    env_done(env);
    return n;
}

/* SparseOps (Bindings.cpp:165-182). The SparseArrayState the reference constructs is flattened into caller buffers:
 * out_counts = {count, n_offsets, n_indirections}; returns 0, 1 on a Java exception, 2 when a buffer is too small. */
static int sparse_unpack(JNIEnv *env, jobject state, int kind, void *values, int cap_values, int *indices,
        int *offsets, int cap_offsets, int *indirections, int cap_indirections, int *out_counts) {
    if (env->pending || !state) {
        return env_done(env) ? 1 : 1;
    }
    jobject *f = (jobject *) state->data;
    const int count = f[1]->len, noff = f[2]->len, nind = f[3]->len;
    out_counts[0] = count;
    out_counts[1] = noff;
    out_counts[2] = nind;
    int rc = 0;
    if (f[0]->len != count || count > cap_values || noff > cap_offsets || nind > cap_indirections) {
        rc = 2;
    } else {
        memcpy(values, f[0]->data, (size_t) count * (kind == SHIM_INT_ARRAY ? 4 : 8));
        memcpy(indices, f[1]->data, (size_t) count * 4);
        memcpy(offsets, f[2]->data, (size_t) noff * 4);
        memcpy(indirections, f[3]->data, (size_t) nind * 4);
    }
    for (int i = 0; i < 4; i++) {
        free(f[i]->data);
        free(f[i]);
    }
    free(f);
    free(state);
    env_done(env);
    return rc;
}

int sstref_insert_sparse(int kind, const void *oldV, int nold, const int *oldD, const int *oldS, const int *oldDo, int nd,
        const int *oldI, const void *newV, int nnew, int *newI, void *values, int cap_values, int *indices, int *offsets,
        int cap_offsets, int *indirections, int cap_indirections, int *out_counts) {
    JNIEnv *env = env_get();
    _jobject ov = arr(kind, oldV, nold), od = IARR(oldD, nd), os = IARR(oldS, nd), odo = IARR(oldDo, nd + 1),
             oi = IARR(oldI, nold), nv = arr(kind, newV, nnew), ni = IARR(newI, nnew);
    jobject state = SparseOps::insert(env, 0, &ov, &od, &os, &odo, &oi, &nv, &ni);
    return sparse_unpack(env, state, kind, values, cap_values, indices, offsets, cap_offsets, indirections,
            cap_indirections, out_counts);
}

int sstref_slice_sparse(int kind, const int *slices, int nslices3, const void *srcV, int nsrc, const int *srcD,
        const int *srcS, const int *srcDo, const int *srcI, const int *srcIo, int nsrcIo, const int *srcIi,
        const void *dstV, int ndst, const int *dstD, const int *dstS, const int *dstDo, const int *dstI, const int *dstIo,
        int ndstIo, const int *dstIi, int nd, void *values, int cap_values, int *indices, int *offsets, int cap_offsets,
        int *indirections, int cap_indirections, int *out_counts) {
    JNIEnv *env = env_get();
    _jobject sl = IARR(slices, nslices3), sv = arr(kind, srcV, nsrc), sd = IARR(srcD, nd), ss = IARR(srcS, nd),
             sdo = IARR(srcDo, nd + 1), si = IARR(srcI, nsrc), sio = IARR(srcIo, nsrcIo), sii = IARR(srcIi, nd * nsrc),
             dv = arr(kind, dstV, ndst), dd = IARR(dstD, nd), ds = IARR(dstS, nd), ddo = IARR(dstDo, nd + 1),
             di = IARR(dstI, ndst), dio = IARR(dstIo, ndstIo), dii = IARR(dstIi, nd * ndst);
    jobject state = SparseOps::slice(env, 0, &sl, &sv, &sd, &ss, &sdo, &si, &sio, &sii, &dv, &dd, &ds, &ddo, &di, &dio,
            &dii);
    return sparse_unpack(env, state, kind, values, cap_values, indices, offsets, cap_offsets, indirections,
            cap_indirections, out_counts);
}

/* NativeImageKernel (Bindings.cpp:144-158) */
int sstref_integral_image(const double *src, int nsrc, const int *sD, const int *sS, double *dst, int ndst,
        const int *dD, const int *dS, int nd) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), d = DARR(dst, ndst), dd = IARR(dD, nd),
             ds = IARR(dS, nd);
    NativeImageKernel::createIntegralImage(env, 0, &s, &sd, &ss, &d, &dd, &ds);
    return env_done(env);
}

int sstref_integral_histogram(const double *src, int nsrc, const int *sD, const int *sS, int nd, const int *mem,
        int nmem, double *dst, int ndst, const int *dD, const int *dS) {
    JNIEnv *env = env_get();
    _jobject s = DARR(src, nsrc), sd = IARR(sD, nd), ss = IARR(sS, nd), m = IARR(mem, nmem), d = DARR(dst, ndst),
             dd = IARR(dD, nd + 1), ds = IARR(dS, nd + 1);
    NativeImageKernel::createIntegralHistogram(env, 0, &s, &sd, &ss, &m, &d, &dd, &ds);
    return env_done(env);
}

}  // extern "C"

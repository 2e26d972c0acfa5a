// jni_glue.cpp -- the JNI face of libsst_cuda.so: one export per native method of the Java providers
// org.sharedx.cuda.CudaArrayKernel / CudaImageKernel / CudaFftService (shared_b200/java), each a pin -> host-tier C ABI call ->
// release -> throw, plus the JNI_OnLoad that registers both providers.
//
// It mirrors, symbol for symbol, what the reference's libsst / libsstx export for their providers:
//   native/src/shared/Bindings.cpp:26-142        Java_org_shared_array_jni_NativeArrayKernel_*
//   native/src/shared/Bindings.cpp:144-158       Java_org_shared_image_jni_NativeImageKernel_*
//   native/src/sharedx/BindingsX.cpp:26-29       Java_org_sharedx_fftw_Plan_transform
//   native/src/shared_common/Library.cpp:33-126  JNI_OnLoad / JNI_OnUnload, Services.registerService
// Ownership follows native/include/shared/Common.hpp:134-145: every array is pinned with
// GetPrimitiveArrayCritical, inputs are released with JNI_ABORT (no copy-back), outputs with mode 0, and no
// JNI call is made while anything is pinned. Errors become java.lang.RuntimeException carrying
// sstc_last_error(), unless an exception is already pending (Common.cpp:173-178).
//
// Built only where a JDK's <jni.h> exists; `make jni_check` syntax-checks it against jni/stub/jni.h elsewhere.
#include <jni.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../../include/sst_cuda.h"

namespace {

jclass g_double_array = nullptr, g_int_array = nullptr, g_object_array = nullptr;
jclass g_library = nullptr, g_services = nullptr;

struct Pin {
    JNIEnv *env;
    jarray arr;
    void *ptr;
    jint mode;
    Pin(JNIEnv *e, jarray a, bool read_only)
        : env(e), arr(a), ptr(a ? e->GetPrimitiveArrayCritical(a, nullptr) : nullptr), mode(read_only ? JNI_ABORT : 0) {}
    ~Pin() {
        if (ptr) {
            env->ReleasePrimitiveArrayCritical(arr, ptr, mode);
        }
    }
    template <class T>
    T *as() const {
        return static_cast<T *>(ptr);
    }
};

void raise(JNIEnv *env, const char *msg) {
    if (!env->ExceptionCheck()) {
        jclass c = env->FindClass("java/lang/RuntimeException");
        if (c) {
            env->ThrowNew(c, msg);
        }
    }
}

void raise_last(JNIEnv *env, int rc) {
    if (rc != 0) {
        raise(env, sstc_last_error());
    }
}

jsize len(JNIEnv *env, jarray a) { return a ? env->GetArrayLength(a) : 0; }

enum Kind { K_DOUBLE, K_INT, K_OBJECT, K_BAD };

// NativeArrayKernel::getArrayType (native/src/shared/NativeArrayKernel.cpp:115-136)
Kind pair_kind(JNIEnv *env, jobject a, jobject b) {
    if (a && b && env->IsInstanceOf(a, g_double_array) && env->IsInstanceOf(b, g_double_array)) {
        return K_DOUBLE;
    }
    if (a && b && env->IsInstanceOf(a, g_int_array) && env->IsInstanceOf(b, g_int_array)) {
        return K_INT;
    }
    if (a && b && env->IsInstanceOf(a, g_object_array) && env->IsInstanceOf(b, g_object_array)) {
        return K_OBJECT;
    }
    return K_BAD;
}

// Object[] map / slice: references cannot live in HBM, so they are moved on the JVM side. Only the index
// arithmetic is done here (no array contents are computed on the CPU).
bool dims_ok(const std::vector<jint> &d, const std::vector<jint> &s, jlong n) {
    jlong acc = 0;
    for (size_t i = 0; i < d.size(); i++) {
        if (d[i] < 0 || s[i] < 0) {
            return false;
        }
        acc += (jlong) (d[i] - 1) * s[i];
    }
    return acc == n - 1;
}

std::vector<jint> copy_ints(JNIEnv *env, jintArray a) {
    std::vector<jint> v((size_t) len(env, a));
    if (!v.empty()) {
        Pin p(env, a, true);
        memcpy(v.data(), p.ptr, v.size() * sizeof(jint));
    }
    return v;
}

void move_objects(JNIEnv *env, bool is_map, jintArray spec, jobjectArray src, jintArray srcD, jintArray srcS,
        jobjectArray dst, jintArray dstD, jintArray dstS) {
    if (!spec || !srcD || !srcS || !dstD || !dstS) {
        return raise(env, "Invalid arguments");
    }
    const std::vector<jint> sp = copy_ints(env, spec), sD = copy_ints(env, srcD), sS = copy_ints(env, srcS),
                            dD = copy_ints(env, dstD), dS = copy_ints(env, dstS);
    const size_t nd = sD.size();
    if (sS.size() != nd || dD.size() != nd || dS.size() != nd || (is_map ? sp.size() != 3 * nd : sp.size() % 3 != 0)) {
        return raise(env, "Invalid arguments");
    }
    const jlong ns = len(env, src), ndst = len(env, dst);
    if (!dims_ok(sD, sS, ns) || !dims_ok(dD, dS, ndst)) {
        return raise(env, "Invalid dimensions and/or strides");
    }
    // per-dimension (source index, destination index) lists
    std::vector<std::vector<std::pair<jint, jint>>> lists(nd);
    if (is_map) {
        for (size_t d = 0; d < nd; d++) {
            if (sp[3 * d + 2] < 0) {
                return raise(env, "Invalid mapping parameters");
            }
        }
        if (ns == 0 || ndst == 0) {
            return;
        }
        for (size_t d = 0; d < nd; d++) {
            for (jint j = 0; j < sp[3 * d + 2]; j++) {
                const jint si = (((sp[3 * d] + j) % sD[d]) + sD[d]) % sD[d];
                const jint di = (((sp[3 * d + 1] + j) % dD[d]) + dD[d]) % dD[d];
                lists[d].push_back({si, di});
            }
        }
    } else {
        for (size_t i = 0; i + 2 < sp.size(); i += 3) {
            const jint si = sp[i], di = sp[i + 1], dim = sp[i + 2];
            if (!(dim >= 0 && (size_t) dim < nd)) {
                return raise(env, "Invalid dimension");
            }
            if (!(si >= 0 && si < sD[(size_t) dim]) || !(di >= 0 && di < dD[(size_t) dim])) {
                return raise(env, "Invalid index");
            }
            lists[(size_t) dim].push_back({si, di});
        }
    }
    jlong total = 1;
    for (size_t d = 0; d < nd; d++) {
        total *= (jlong) lists[d].size();
    }
    std::vector<size_t> idx(nd, 0);
    for (jlong m = 0; m < total; m++) {  // row-major order, last dimension fastest: later writes win
        jlong so = 0, dof = 0;
        for (size_t d = 0; d < nd; d++) {
            so += (jlong) lists[d][idx[d]].first * sS[d];
            dof += (jlong) lists[d][idx[d]].second * dS[d];
        }
        jobject ref = env->GetObjectArrayElement(src, (jsize) so);
        env->SetObjectArrayElement(dst, (jsize) dof, ref);
        env->DeleteLocalRef(ref);  // O(total) local references otherwise: the JVM only guarantees 16 per native frame
        for (size_t d = nd; d-- > 0;) {
            if (++idx[d] < lists[d].size()) {
                break;
            }
            idx[d] = 0;
        }
    }
}

void move_any(JNIEnv *env, bool is_map, jintArray spec, jobject srcV, jintArray srcD, jintArray srcS, jobject dstV,
        jintArray dstD, jintArray dstS) {
    if (!spec || !srcV || !srcD || !srcS || !dstV || !dstD || !dstS) {
        return raise(env, "Invalid arguments");
    }
    const Kind kind = pair_kind(env, srcV, dstV);
    if (kind == K_BAD) {
        return raise(env, "Invalid array types");  // NativeArrayKernel.cpp:134, ObjectArrayTest.java:301
    }
    if (kind == K_OBJECT) {
        return move_objects(env, is_map, spec, (jobjectArray) srcV, srcD, srcS, (jobjectArray) dstV, dstD, dstS);
    }
    const jsize nsp = len(env, spec), nd = len(env, srcD);
    if (len(env, srcS) != nd || len(env, dstD) != nd || len(env, dstS) != nd) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, (jarray) srcV), ndst = len(env, (jarray) dstV);
    const int eb = kind == K_DOUBLE ? 8 : 4;
    int rc;
    {
        Pin sp(env, spec, true), s(env, (jarray) srcV, true), sD(env, srcD, true), sS(env, srcS, true),
                d(env, (jarray) dstV, false), dD(env, dstD, true), dS(env, dstS, true);
        rc = is_map ? sstc_map(eb, sp.as<int32_t>(), nsp, s.ptr, ns, sD.as<int32_t>(), sS.as<int32_t>(), d.ptr, ndst,
                              dD.as<int32_t>(), dS.as<int32_t>(), nd)
                    : sstc_slice(eb, sp.as<int32_t>(), nsp, s.ptr, ns, sD.as<int32_t>(), sS.as<int32_t>(), d.ptr, ndst,
                              dD.as<int32_t>(), dS.as<int32_t>(), nd);
    }
    raise_last(env, rc);
}

jclass global_class(JNIEnv *env, const char *name) {
    jclass local = env->FindClass(name);
    return local ? (jclass) env->NewGlobalRef(local) : nullptr;
}

bool register_service(JNIEnv *env, jmethodID method, const char *spec, const char *impl) {
    jclass s = env->FindClass(spec), i = env->FindClass(impl);
    if (!s || !i) {
        return false;
    }
    env->CallStaticVoidMethod(g_services, method, s, i);
    return !env->ExceptionCheck();
}

}  // namespace

#define AK(name) Java_org_sharedx_cuda_CudaArrayKernel_##name

extern "C" {

// ---- library lifecycle (Library.cpp:33-126) ---------------------------------------------------------------

JNIEXPORT jint JNICALL JNI_OnLoad(JavaVM *jvm, void *) {
    void *p = nullptr;
    if (jvm->GetEnv(&p, JNI_VERSION_1_4) != JNI_OK) {
        return JNI_ERR;
    }
    JNIEnv *env = static_cast<JNIEnv *>(p);
    g_double_array = global_class(env, "[D");
    g_int_array = global_class(env, "[I");
    g_object_array = global_class(env, "[Ljava/lang/Object;");
    g_library = global_class(env, "org/shared/metaclass/Library");
    g_services = global_class(env, "org/shared/util/Services");
    if (!g_double_array || !g_int_array || !g_object_array || !g_library || !g_services) {
        return JNI_VERSION_1_4;  // a ClassNotFoundException is pending; Library.initialized stays false
    }
    const char *dev = getenv("SSTC_DEVICE");
    if (sstc_init(dev ? atoi(dev) : 0) != 0) {
        raise(env, sstc_last_error());  // no CUDA device: the providers are not registered (no CPU fallback)
        return JNI_VERSION_1_4;
    }
    jmethodID reg = env->GetStaticMethodID(g_services, "registerService", "(Ljava/lang/Class;Ljava/lang/Class;)V");
    jfieldID init = env->GetStaticFieldID(g_library, "initialized", "Z");
    if (reg && init &&
            register_service(env, reg, "org/shared/array/kernel/ArrayKernel", "org/sharedx/cuda/CudaArrayKernel") &&
            register_service(env, reg, "org/shared/fft/FftService", "org/sharedx/cuda/CudaFftService") &&
            register_service(env, reg, "org/shared/image/kernel/ImageKernel", "org/sharedx/cuda/CudaImageKernel")) {
        env->SetStaticBooleanField(g_library, init, JNI_TRUE);
    }
    return JNI_VERSION_1_4;
}

JNIEXPORT void JNICALL JNI_OnUnload(JavaVM *jvm, void *) {
    void *p = nullptr;
    if (jvm->GetEnv(&p, JNI_VERSION_1_4) != JNI_OK) {
        return;
    }
    JNIEnv *env = static_cast<JNIEnv *>(p);
    if (g_library) {
        jfieldID init = env->GetStaticFieldID(g_library, "initialized", "Z");
        if (init) {
            env->SetStaticBooleanField(g_library, init, JNI_FALSE);
        }
    }
    jclass *all[] = {&g_double_array, &g_int_array, &g_object_array, &g_library, &g_services};
    for (jclass *c : all) {
        if (*c) {
            env->DeleteGlobalRef(*c);
            *c = nullptr;
        }
    }
    sstc_shutdown();
}

// ---- FftService (BindingsX.cpp:26-29 -> Plan::transform, Plan.cpp:44-99) --------------------------------------

JNIEXPORT void JNICALL Java_org_sharedx_cuda_CudaFftService_transform(JNIEnv *env, jobject, jint kind, jintArray dims,
        jdoubleArray in, jdoubleArray out) {
    if (!dims || !in || !out) {
        return raise(env, "Invalid arguments");
    }
    const jsize nd = len(env, dims), nin = len(env, in), nout = len(env, out);
    int rc;
    {
        Pin d(env, dims, true), i(env, in, true), o(env, out, false);
        rc = sstc_fft(kind, nd, d.as<int32_t>(), i.as<double>(), nin, o.as<double>(), nout);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL Java_org_sharedx_cuda_CudaFftService_setDevices(JNIEnv *env, jobject, jintArray devices) {
    if (!devices) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, devices);
    int rc;
    {
        Pin d(env, devices, true);
        rc = sstc_fft_set_devices(n, d.as<int32_t>());
    }
    raise_last(env, rc);
}

// ---- CudaComplexArray: the device-resident twin; static natives, device tier of the C ABI ----------
// (the class parameter is declared jobject: jclass is a jobject, and the repository's consistency tests read one shape)

#define CA(name) Java_org_sharedx_cuda_CudaComplexArray_##name

JNIEXPORT jlong JNICALL CA(allocate)(JNIEnv *env, jobject, jint nDoubles) {
    void *p = nullptr;
    int rc = nDoubles < 0 ? 1 : sstc_malloc(&p, (int64_t) nDoubles * 8);
    if (rc == 0 && p) {
        rc = sstc_memset(p, 0, (int64_t) nDoubles * 8, nullptr);
    }
    if (nDoubles < 0) {
        raise(env, "Invalid arguments");
        return 0;
    }
    raise_last(env, rc);
    return (jlong) (intptr_t) p;
}

JNIEXPORT void JNICALL CA(free)(JNIEnv *env, jobject, jlong pointer) {
    raise_last(env, sstc_free((void *) (intptr_t) pointer));
}

JNIEXPORT void JNICALL CA(upload)(JNIEnv *env, jobject, jlong pointer, jdoubleArray values) {
    if (!values || !pointer) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, values);
    int rc;
    {
        Pin v(env, values, true);
        rc = sstc_h2d((void *) (intptr_t) pointer, v.as<double>(), (int64_t) n * 8, nullptr);
        if (rc == 0) {
            rc = sstc_stream_sync(nullptr);  // the pinned JVM array must not move before the copy has read it
        }
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL CA(download)(JNIEnv *env, jobject, jlong pointer, jdoubleArray values) {
    if (!values || !pointer) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, values);
    int rc;
    {
        Pin v(env, values, false);
        rc = sstc_d2h(v.as<double>(), (const void *) (intptr_t) pointer, (int64_t) n * 8, nullptr);
        if (rc == 0) {
            rc = sstc_stream_sync(nullptr);
        }
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL CA(eMul)(JNIEnv *env, jobject, jlong lhs, jlong rhs, jlong dst, jint nDoubles) {
    // CE_MUL = 2 (ArrayKernel.java:247), complex pairs
    raise_last(env, sstc_eop_dev(2, SSTC_ELEM_COMPLEX, (const void *) (intptr_t) lhs, (const void *) (intptr_t) rhs,
                            (void *) (intptr_t) dst, nDoubles, nullptr));
}

JNIEXPORT void JNICALL CA(transform)(JNIEnv *env, jobject, jint kind, jintArray dims, jlong in, jlong out) {
    if (!dims || !in || !out) {
        return raise(env, "Invalid arguments");
    }
    // plans are cached per (kind, dims), like FftwService's plan cache (FftwService.java:58-67,221-236)
    static std::mutex mutex;
    static std::map<std::vector<int32_t>, sstc_fft_plan> plans;
    const jsize nd = len(env, dims);
    std::vector<int32_t> key(1, kind);
    {
        Pin d(env, dims, true);
        key.insert(key.end(), d.as<int32_t>(), d.as<int32_t>() + nd);
    }
    int rc = 0;
    sstc_fft_plan plan = nullptr;
    {
        std::lock_guard<std::mutex> lock(mutex);
        auto it = plans.find(key);
        if (it == plans.end()) {
            rc = sstc_fft_plan_create(&plan, kind, nd, key.data() + 1);
            if (rc == 0) {
                plans[key] = plan;
            }
        } else {
            plan = it->second;
        }
    }
    if (rc == 0) {
        rc = sstc_fft_plan_execute(plan, (const double *) (intptr_t) in, (double *) (intptr_t) out, nullptr);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL CA(fftShift)(JNIEnv *env, jobject, jlong src, jlong dst, jintArray dims, jboolean inverse) {
    if (!dims || !src || !dst) {
        return raise(env, "Invalid arguments");
    }
    const jsize nd = len(env, dims);
    std::vector<int32_t> d;
    {
        Pin p(env, dims, true);
        d.assign(p.as<int32_t>(), p.as<int32_t>() + nd);
    }
    raise_last(env, sstc_fftshift_dev((const double *) (intptr_t) src, (double *) (intptr_t) dst, d.data(), nd, inverse ? 1 : 0,
                            nullptr));
}

JNIEXPORT void JNICALL CA(synchronize)(JNIEnv *env, jobject) { raise_last(env, sstc_stream_sync(nullptr)); }

// ---- ArrayKernel (Bindings.cpp:26-119) ---------------------------------------------------------------------------

JNIEXPORT void JNICALL AK(randomize)(JNIEnv *env, jobject) { raise_last(env, sstc_srand((uint64_t) time(nullptr))); }

JNIEXPORT void JNICALL AK(derandomize)(JNIEnv *env, jobject) { raise_last(env, sstc_srand(0xdeadbeefcacabeadULL)); }

JNIEXPORT void JNICALL AK(map)(JNIEnv *env, jobject, jintArray bounds, jobject srcV, jintArray srcD, jintArray srcS,
        jobject dstV, jintArray dstD, jintArray dstS) {
    move_any(env, true, bounds, srcV, srcD, srcS, dstV, dstD, dstS);
}

JNIEXPORT void JNICALL AK(slice)(JNIEnv *env, jobject, jintArray slices, jobject srcV, jintArray srcD, jintArray srcS,
        jobject dstV, jintArray dstD, jintArray dstS) {
    move_any(env, false, slices, srcV, srcD, srcS, dstV, dstD, dstS);
}

JNIEXPORT void JNICALL AK(rrOp)(JNIEnv *env, jobject, jint type, jdoubleArray srcV, jintArray srcD, jintArray srcS,
        jdoubleArray dstV, jintArray dstD, jintArray dstS, jintArray opDims) {
    if (!srcV || !srcD || !srcS || !dstV || !dstD || !dstS || !opDims) {
        return raise(env, "Invalid arguments");
    }
    const jsize nd = len(env, srcD), ns = len(env, srcV), ndst = len(env, dstV), nop = len(env, opDims);
    if (len(env, srcS) != nd || len(env, dstD) != nd || len(env, dstS) != nd) {
        return raise(env, "Invalid arguments");
    }
    int rc;
    {
        Pin s(env, srcV, true), sD(env, srcD, true), sS(env, srcS, true), d(env, dstV, false), dD(env, dstD, true),
                dS(env, dstS, true), od(env, opDims, false);  // opDims is sorted in place, as the reference does
        rc = sstc_rrop(type, s.as<double>(), ns, sD.as<int32_t>(), sS.as<int32_t>(), d.as<double>(), ndst,
                dD.as<int32_t>(), dS.as<int32_t>(), nd, od.as<int32_t>(), nop);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(riOp)(JNIEnv *env, jobject, jint type, jdoubleArray srcV, jintArray srcD, jintArray srcS,
        jintArray dstV, jint dim) {
    if (!srcV || !srcD || !srcS || !dstV) {
        return raise(env, "Invalid arguments");
    }
    const jsize nd = len(env, srcD), ns = len(env, srcV), ndst = len(env, dstV);
    if (len(env, srcS) != nd) {
        return raise(env, "Invalid arguments");
    }
    int rc;
    {
        Pin s(env, srcV, false), sD(env, srcD, true), sS(env, srcS, true), d(env, dstV, false);  // RI_SORT writes src
        rc = sstc_riop(type, s.as<double>(), ns, sD.as<int32_t>(), sS.as<int32_t>(), nd, d.as<int32_t>(), ndst, dim);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(rdOp)(JNIEnv *env, jobject, jint type, jdoubleArray srcV, jintArray srcD, jintArray srcS,
        jdoubleArray dstV, jintArray opDims) {
    if (!srcV || !srcD || !srcS || !dstV || !opDims) {
        return raise(env, "Invalid arguments");
    }
    const jsize nd = len(env, srcD), ns = len(env, srcV), ndst = len(env, dstV), nop = len(env, opDims);
    if (len(env, srcS) != nd) {
        return raise(env, "Invalid arguments");
    }
    int rc;
    {
        Pin s(env, srcV, true), sD(env, srcD, true), sS(env, srcS, true), d(env, dstV, false), od(env, opDims, true);
        rc = sstc_rdop(type, s.as<double>(), ns, sD.as<int32_t>(), sS.as<int32_t>(), nd, d.as<double>(), ndst,
                od.as<int32_t>(), nop);
    }
    raise_last(env, rc);
}

JNIEXPORT jdouble JNICALL AK(raOp)(JNIEnv *env, jobject, jint type, jdoubleArray srcV) {
    if (!srcV) {
        raise(env, "Invalid arguments");
        return 0.0;
    }
    const jsize n = len(env, srcV);
    double out = 0.0;
    int rc;
    {
        Pin s(env, srcV, true);
        rc = sstc_raop(type, s.as<double>(), n, &out);
    }
    raise_last(env, rc);
    return out;
}

JNIEXPORT jdoubleArray JNICALL AK(caOp)(JNIEnv *env, jobject, jint type, jdoubleArray srcV) {
    if (!srcV) {
        raise(env, "Invalid arguments");
        return nullptr;
    }
    const jsize n = len(env, srcV);
    double out[2] = {0.0, 0.0};
    int rc;
    {
        Pin s(env, srcV, true);
        rc = sstc_caop(type, s.as<double>(), n, out);
    }
    if (rc != 0) {
        raise_last(env, rc);
        return nullptr;
    }
    jdoubleArray res = env->NewDoubleArray(2);  // ElementOps.cpp:98-103
    if (res) {
        env->SetDoubleArrayRegion(res, 0, 2, out);
    }
    return res;
}

JNIEXPORT void JNICALL AK(ruOp)(JNIEnv *env, jobject, jint type, jdouble a, jdoubleArray srcV) {
    if (!srcV) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, srcV);
    int rc;
    {
        Pin s(env, srcV, false);
        rc = sstc_ruop(type, a, s.as<double>(), n);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(cuOp)(JNIEnv *env, jobject, jint type, jdouble aRe, jdouble aIm, jdoubleArray srcV) {
    if (!srcV) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, srcV);
    int rc;
    {
        Pin s(env, srcV, false);
        rc = sstc_cuop(type, aRe, aIm, s.as<double>(), n);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(iuOp)(JNIEnv *env, jobject, jint type, jint a, jintArray srcV) {
    if (!srcV) {
        return raise(env, "Invalid arguments");
    }
    const jsize n = len(env, srcV);
    int rc;
    {
        Pin s(env, srcV, false);
        rc = sstc_iuop(type, a, s.as<int32_t>(), n);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(eOp)(JNIEnv *env, jobject, jint type, jobject lhsV, jobject rhsV, jobject dstV,
        jboolean complex) {
    // ElementOps.cpp:285-321: three double[] (real or complex by the flag) or three int[]
    int elem;
    if (pair_kind(env, lhsV, rhsV) == K_DOUBLE && pair_kind(env, rhsV, dstV) == K_DOUBLE) {
        elem = complex ? SSTC_ELEM_COMPLEX : SSTC_ELEM_REAL;
    } else if (pair_kind(env, lhsV, rhsV) == K_INT && pair_kind(env, rhsV, dstV) == K_INT) {
        elem = SSTC_ELEM_INT;
    } else {
        return raise(env, "Invalid array types");
    }
    const jsize n = len(env, (jarray) dstV);
    if (len(env, (jarray) lhsV) != n || len(env, (jarray) rhsV) != n) {
        return raise(env, "Invalid array lengths");
    }
    int rc;
    {
        Pin l(env, (jarray) lhsV, true), r(env, (jarray) rhsV, true), d(env, (jarray) dstV, false);
        rc = sstc_eop(type, elem, l.ptr, r.ptr, d.ptr, n);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(convert)(JNIEnv *env, jobject, jint type, jobject srcV, jboolean isSrcComplex, jobject dstV,
        jboolean isDstComplex) {
    if (!srcV || !dstV) {
        return raise(env, "Invalid arguments");
    }
    const bool sd = env->IsInstanceOf(srcV, g_double_array), si = env->IsInstanceOf(srcV, g_int_array),
               dd = env->IsInstanceOf(dstV, g_double_array);
    int conv;
    if (sd && !isSrcComplex && dd && isDstComplex) {  // ElementOps.cpp:329-401
        conv = SSTC_CONV_R_TO_C;
    } else if (sd && isSrcComplex && dd && !isDstComplex) {
        conv = SSTC_CONV_C_TO_R;
    } else if (si && !isSrcComplex && dd && !isDstComplex) {
        conv = SSTC_CONV_I_TO_R;
    } else {
        return raise(env, "Invalid array types");
    }
    const jsize ns = len(env, (jarray) srcV), nd = len(env, (jarray) dstV);
    int rc;
    {
        Pin s(env, (jarray) srcV, true), d(env, (jarray) dstV, false);
        rc = sstc_convert(type, conv, s.ptr, ns, d.as<double>(), nd);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(mul)(JNIEnv *env, jobject, jdoubleArray lhsV, jdoubleArray rhsV, jint lr, jint rc_,
        jdoubleArray dstV, jboolean complex) {
    if (!lhsV || !rhsV || !dstV) {
        return raise(env, "Invalid arguments");
    }
    const jsize nl = len(env, lhsV), nr = len(env, rhsV), nd = len(env, dstV);
    int rc;
    {
        Pin l(env, lhsV, true), r(env, rhsV, true), d(env, dstV, false);
        rc = sstc_mul(l.as<double>(), nl, r.as<double>(), nr, lr, rc_, d.as<double>(), nd, complex ? 1 : 0);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(diag)(JNIEnv *env, jobject, jdoubleArray srcV, jdoubleArray dstV, jint size,
        jboolean complex) {
    if (!srcV || !dstV) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nd = len(env, dstV);
    int rc;
    {
        Pin s(env, srcV, true), d(env, dstV, false);
        rc = sstc_diag(s.as<double>(), ns, d.as<double>(), nd, size, complex ? 1 : 0);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(invert)(JNIEnv *env, jobject, jdoubleArray srcV, jdoubleArray dstV, jint size) {
    if (!srcV || !dstV) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nd = len(env, dstV);
    int rc;
    {
        Pin s(env, srcV, true), d(env, dstV, false);
        rc = sstc_invert(s.as<double>(), ns, d.as<double>(), nd, size);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(eigs)(JNIEnv *env, jobject, jdoubleArray srcV, jdoubleArray vecV, jdoubleArray valV,
        jint size) {
    if (!srcV || !vecV || !valV) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nvec = len(env, vecV), nval = len(env, valV);
    int rc;
    {
        Pin s(env, srcV, true), ve(env, vecV, false), va(env, valV, false);
        rc = sstc_eigs(s.as<double>(), ns, ve.as<double>(), nvec, va.as<double>(), nval, size);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL AK(svd)(JNIEnv *env, jobject, jdoubleArray srcV, jint srcStrideRow, jint srcStrideCol,
        jdoubleArray uV, jdoubleArray sV, jdoubleArray vV, jint nRows, jint nCols) {
    if (!srcV || !uV || !sV || !vV) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nu = len(env, uV), nsv = len(env, sV), nv = len(env, vV);
    int rc;
    {
        Pin s(env, srcV, true), u(env, uV, false), sv(env, sV, false), v(env, vV, false);
        rc = sstc_svd(s.as<double>(), ns, srcStrideRow, srcStrideCol, u.as<double>(), nu, sv.as<double>(), nsv,
                v.as<double>(), nv, nRows, nCols);
    }
    raise_last(env, rc);
}

JNIEXPORT jintArray JNICALL AK(find)(JNIEnv *env, jobject, jintArray srcV, jintArray srcD, jintArray srcS,
        jintArray logical) {
    if (!srcV || !srcD || !srcS || !logical) {
        raise(env, "Invalid arguments");
        return nullptr;
    }
    const jsize ns = len(env, srcV), rank = len(env, srcD);
    if (rank != len(env, srcS) || rank != len(env, logical)) {
        raise(env, "Invalid arguments");
        return nullptr;
    }
    std::vector<int32_t> line((size_t) (ns > 0 ? ns : 1));  // a line is never longer than the array
    int32_t count = 0;
    int rc;
    {
        Pin s(env, srcV, true), sd(env, srcD, true), ss(env, srcS, true), lg(env, logical, true);
        rc = sstc_find(s.as<int32_t>(), ns, sd.as<int32_t>(), ss.as<int32_t>(), rank, lg.as<int32_t>(), line.data(),
                (int64_t) line.size(), &count);
    }
    if (rc != 0) {
        raise_last(env, rc);
        return nullptr;
    }
    jintArray res = env->NewIntArray(count);  // IndexOps.cpp:38 (Common::newIntArray)
    if (res && count > 0) {
        env->SetIntArrayRegion(res, 0, count, reinterpret_cast<const jint *>(line.data()));
    }
    return res;
}

// ---- sparse arrays (Bindings.cpp:165-182 -> SparseOpsInsert.cpp / SparseOpsSlice.cpp) -------------------------------
//
// double[] / int[] values go through the C ABI as they are. Object[] values cannot live in HBM: the same call is made
// with POSITIONS as 4-byte values (old / destination: 0 .., new / source: n_old ..), and the references are gathered
// into the result array here -- all index work still happens on the device.

namespace {

// new SparseArrayState(values, indices, indirectionOffsets, indirections): NativeArrayKernel.cpp:103-115
jobject new_sparse_state(JNIEnv *env, const sstc_sparse_state &st, Kind kind, jobjectArray first, jsize n_first,
        jobjectArray second) {
    jobject values = nullptr;
    if (kind == K_DOUBLE) {
        jdoubleArray a = env->NewDoubleArray(st.count);
        if (a && st.count > 0) {
            env->SetDoubleArrayRegion(a, 0, st.count, static_cast<const jdouble *>(st.values));
        }
        values = a;
    } else if (kind == K_INT) {
        jintArray a = env->NewIntArray(st.count);
        if (a && st.count > 0) {
            env->SetIntArrayRegion(a, 0, st.count, static_cast<const jint *>(st.values));
        }
        values = a;
    } else {
        // component type of the first array (SparseOps.cpp:96-98: getComponentType(GetObjectClass(oldV)))
        jclass cls = env->FindClass("java/lang/Class");
        jmethodID get = cls ? env->GetMethodID(cls, "getComponentType", "()Ljava/lang/Class;") : nullptr;
        jclass comp = get ? static_cast<jclass>(env->CallObjectMethod(env->GetObjectClass(first), get)) : nullptr;
        jobjectArray a = comp ? env->NewObjectArray(st.count, comp, nullptr) : nullptr;
        const int32_t *pos = static_cast<const int32_t *>(st.values);
        for (jsize i = 0; a && i < st.count; i++) {
            jobject ref = pos[i] < n_first ? env->GetObjectArrayElement(first, pos[i])
                                           : env->GetObjectArrayElement(second, pos[i] - n_first);
            env->SetObjectArrayElement(a, i, ref);
            env->DeleteLocalRef(ref);
        }
        values = a;
    }
    jintArray indices = env->NewIntArray(st.count);
    jintArray offsets = env->NewIntArray(st.n_offsets);
    jintArray indirections = env->NewIntArray(st.n_indirections);
    if (!values || !indices || !offsets || !indirections || env->ExceptionCheck()) {
        return nullptr;  // OutOfMemoryError pending
    }
    if (st.count > 0) {
        env->SetIntArrayRegion(indices, 0, st.count, reinterpret_cast<const jint *>(st.indices));
    }
    if (st.n_offsets > 0) {
        env->SetIntArrayRegion(offsets, 0, st.n_offsets, reinterpret_cast<const jint *>(st.indirection_offsets));
    }
    if (st.n_indirections > 0) {
        env->SetIntArrayRegion(indirections, 0, st.n_indirections, reinterpret_cast<const jint *>(st.indirections));
    }
    jclass cls = env->FindClass("org/shared/array/sparse/SparseArrayState");
    jmethodID ctor = cls ? env->GetMethodID(cls, "<init>", "(Ljava/lang/Object;[I[I[I)V") : nullptr;
    return ctor ? env->NewObject(cls, ctor, values, indices, offsets, indirections) : nullptr;
}

std::vector<int32_t> positions(jsize n, int32_t base) {
    std::vector<int32_t> v((size_t) (n > 0 ? n : 1));
    for (jsize i = 0; i < n; i++) {
        v[(size_t) i] = base + i;
    }
    return v;
}

}  // namespace

JNIEXPORT jobject JNICALL AK(insertSparse)(JNIEnv *env, jobject, jobject oldV, jintArray oldD, jintArray oldS,
        jintArray oldDo, jintArray oldI, jobject newV, jintArray newI) {
    if (!oldV || !oldD || !oldS || !oldDo || !oldI || !newV || !newI) {
        raise(env, "Invalid arguments");
        return nullptr;
    }
    const Kind kind = pair_kind(env, newV, oldV);
    if (kind == K_BAD) {
        raise(env, "Invalid array types");
        return nullptr;
    }
    const jsize nd = len(env, oldD), n_old = len(env, static_cast<jarray>(oldV)), n_new = len(env, static_cast<jarray>(newV));
    if (nd != len(env, oldS) || n_old != len(env, oldI) || n_new != len(env, newI) || nd + 1 != len(env, oldDo)) {
        raise(env, "Invalid arguments");  // SparseOpsInsert.cpp:77-83
        return nullptr;
    }
    sstc_sparse_state st;
    int rc;
    if (kind == K_OBJECT) {
        const std::vector<int32_t> ov = positions(n_old, 0), nv = positions(n_new, n_old);
        Pin d(env, oldD, true), s(env, oldS, true), o(env, oldDo, true), oi(env, oldI, true), ni(env, newI, false);
        rc = sstc_insert_sparse(4, ov.data(), n_old, d.as<int32_t>(), s.as<int32_t>(), o.as<int32_t>(), nd,
                oi.as<int32_t>(), nv.data(), n_new, ni.as<int32_t>(), &st);
    } else {
        Pin ov(env, static_cast<jarray>(oldV), true), nv(env, static_cast<jarray>(newV), true), d(env, oldD, true),
                s(env, oldS, true), o(env, oldDo, true), oi(env, oldI, true), ni(env, newI, false);
        rc = sstc_insert_sparse(kind == K_DOUBLE ? 8 : 4, ov.ptr, n_old, d.as<int32_t>(), s.as<int32_t>(), o.as<int32_t>(),
                nd, oi.as<int32_t>(), nv.ptr, n_new, ni.as<int32_t>(), &st);
    }
    if (rc != 0) {
        raise_last(env, rc);
        return nullptr;
    }
    return new_sparse_state(env, st, kind, static_cast<jobjectArray>(oldV), n_old, static_cast<jobjectArray>(newV));
}

JNIEXPORT jobject JNICALL AK(sliceSparse)(JNIEnv *env, jobject, jintArray slices, jobject srcV, jintArray srcD,
        jintArray srcS, jintArray srcDo, jintArray srcI, jintArray srcIo, jintArray srcIi, jobject dstV, jintArray dstD,
        jintArray dstS, jintArray dstDo, jintArray dstI, jintArray dstIo, jintArray dstIi) {
    if (!slices || !srcV || !srcD || !srcS || !srcDo || !srcI || !srcIo || !srcIi || !dstV || !dstD || !dstS || !dstDo ||
            !dstI || !dstIo || !dstIi) {
        raise(env, "Invalid arguments");
        return nullptr;
    }
    const Kind kind = pair_kind(env, srcV, dstV);
    if (kind == K_BAD) {
        raise(env, "Invalid array types");
        return nullptr;
    }
    const jsize nd = len(env, srcD), n_src = len(env, static_cast<jarray>(srcV)), n_dst = len(env, static_cast<jarray>(dstV));
    const jsize n_sl = len(env, slices);
    jlong sum_s = 0, sum_d = 0;
    if (nd == len(env, dstD)) {
        for (jint v : copy_ints(env, srcD)) {
            sum_s += v;
        }
        for (jint v : copy_ints(env, dstD)) {
            sum_d += v;
        }
    }
    if (nd != len(env, srcS) || nd != len(env, dstD) || nd != len(env, dstS) || n_sl % 3 != 0 || n_src != len(env, srcI) ||
            n_dst != len(env, dstI) || nd + 1 != len(env, srcDo) || (jlong) nd * n_src != len(env, srcIi) ||
            nd + 1 != len(env, dstDo) || (jlong) nd * n_dst != len(env, dstIi) || len(env, srcIo) != sum_s + nd ||
            len(env, dstIo) != sum_d + nd) {
        raise(env, "Invalid arguments");  // SparseOpsSlice.cpp:87-98, 128-131
        return nullptr;
    }
    sstc_sparse_state st;
    int rc;
    {
        // the destination plays "old" (positions 0 ..), the source "new" (positions n_dst ..): SparseOpsSlice.cpp:59
        const std::vector<int32_t> dv = positions(kind == K_OBJECT ? n_dst : 0, 0),
                                   sv = positions(kind == K_OBJECT ? n_src : 0, n_dst);
        Pin sl(env, slices, true), sd(env, srcD, true), ss(env, srcS, true), sdo(env, srcDo, true), si(env, srcI, true),
                sio(env, srcIo, true), sii(env, srcIi, true), dd(env, dstD, true), ds(env, dstS, true), ddo(env, dstDo, true),
                di(env, dstI, true), dio(env, dstIo, true), dii(env, dstIi, true);
        Pin pv(env, kind == K_OBJECT ? nullptr : static_cast<jarray>(srcV), true),
                pd(env, kind == K_OBJECT ? nullptr : static_cast<jarray>(dstV), true);
        rc = sstc_slice_sparse(kind == K_DOUBLE ? 8 : 4, sl.as<int32_t>(), n_sl,
                kind == K_OBJECT ? static_cast<const void *>(sv.data()) : pv.ptr, n_src, sd.as<int32_t>(), ss.as<int32_t>(),
                sdo.as<int32_t>(), si.as<int32_t>(), sio.as<int32_t>(), sii.as<int32_t>(),
                kind == K_OBJECT ? static_cast<const void *>(dv.data()) : pd.ptr, n_dst, dd.as<int32_t>(), ds.as<int32_t>(),
                ddo.as<int32_t>(), di.as<int32_t>(), dio.as<int32_t>(), dii.as<int32_t>(), nd, &st);
    }
    if (rc != 0) {
        raise_last(env, rc);
        return nullptr;
    }
    return new_sparse_state(env, st, kind, static_cast<jobjectArray>(dstV), n_dst, static_cast<jobjectArray>(srcV));
}

// ---- ImageKernel (Bindings.cpp:144-158 -> NativeImageKernel.cpp:32-247) --------------------------------------------

JNIEXPORT void JNICALL Java_org_sharedx_cuda_CudaImageKernel_createIntegralImage(JNIEnv *env, jobject,
        jdoubleArray srcV, jintArray srcD, jintArray srcS, jdoubleArray dstV, jintArray dstD, jintArray dstS) {
    if (!srcV || !srcD || !srcS || !dstV || !dstD || !dstS) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nd = len(env, dstV), rank = len(env, srcD);
    if (rank != len(env, srcS) || rank != len(env, dstD) || rank != len(env, dstS)) {
        return raise(env, "Invalid arguments");
    }
    int rc;
    {
        Pin s(env, srcV, true), sd(env, srcD, true), ss(env, srcS, true), d(env, dstV, false), dd(env, dstD, true),
                ds(env, dstS, true);
        rc = sstc_integral_image(s.as<double>(), ns, sd.as<int32_t>(), ss.as<int32_t>(), d.as<double>(), nd,
                dd.as<int32_t>(), ds.as<int32_t>(), rank);
    }
    raise_last(env, rc);
}

JNIEXPORT void JNICALL Java_org_sharedx_cuda_CudaImageKernel_createIntegralHistogram(JNIEnv *env, jobject,
        jdoubleArray srcV, jintArray srcD, jintArray srcS, jintArray memV, jdoubleArray dstV, jintArray dstD,
        jintArray dstS) {
    if (!srcV || !srcD || !srcS || !memV || !dstV || !dstD || !dstS) {
        return raise(env, "Invalid arguments");
    }
    const jsize ns = len(env, srcV), nm = len(env, memV), nd = len(env, dstV), rank = len(env, srcD);
    if (rank != len(env, srcS) || rank + 1 != len(env, dstD) || rank + 1 != len(env, dstS) || ns != nm) {
        return raise(env, "Invalid arguments");
    }
    int rc;
    {
        Pin s(env, srcV, true), sd(env, srcD, true), ss(env, srcS, true), m(env, memV, true), d(env, dstV, false),
                dd(env, dstD, true), ds(env, dstS, true);
        rc = sstc_integral_histogram(s.as<double>(), ns, sd.as<int32_t>(), ss.as<int32_t>(), rank, m.as<int32_t>(), nm,
                d.as<double>(), nd, dd.as<int32_t>(), ds.as<int32_t>());
    }
    raise_last(env, rc);
}

}  // extern "C"

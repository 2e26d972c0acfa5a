/*
 * TEST INFRASTRUCTURE ONLY -- a minimal stand-in for <jni.h>, written for this repository, so that the
 * reference's native ArrayKernel sources (/root/reference/native/src/shared/*.cpp) compile UNMODIFIED without a
 * JDK and can serve as the parity oracle (oracle/_ref/libsst_ref.so). Java arrays are modelled as
 * {kind, data, len} records owned by the caller (oracle/ref/ref_driver.cpp); only the JNI members those
 * sources use exist. Nothing here is product code and nothing under shared_b200/ includes it.
 */
#ifndef SST_ORACLE_SHIM_JNI_H
#define SST_ORACLE_SHIM_JNI_H

#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#define JNIEXPORT
#define JNICALL
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_ERR (-1)
#define JNI_ABORT 2
#define JNI_COMMIT 1
#define JNI_VERSION_1_4 0x00010004
#define JNI_VERSION_1_6 0x00010006

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef uint16_t jchar;
typedef int16_t jshort;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

enum shim_kind { SHIM_DOUBLE_ARRAY = 1, SHIM_INT_ARRAY = 2, SHIM_OBJECT_ARRAY = 3, SHIM_BYTE_ARRAY = 4,
    SHIM_CLASS = 5, SHIM_OTHER = 6 };

struct _jobject {
    int kind;
    void *data;
    jint len;
    int owned; /* data was allocated by the shim (New*Array) */
};

typedef _jobject *jobject;
typedef jobject jclass;
typedef jobject jthrowable;
typedef jobject jstring;
typedef jobject jarray;
typedef jobject jbooleanArray;
typedef jobject jbyteArray;
typedef jobject jintArray;
typedef jobject jdoubleArray;
typedef jobject jobjectArray;
typedef jobject jweak;

struct _jfieldID {
    int dummy;
};
struct _jmethodID {
    int dummy;
};
typedef _jfieldID *jfieldID;
typedef _jmethodID *jmethodID;

struct JNIEnv_ {
    bool pending;
    std::string message;
    /* what a library's JNI_OnLoad did (tests/jni_harness): Services.registerService(spec, impl) calls by class name,
     * and the last value stored into a static boolean field (Library.initialized) */
    std::vector<std::pair<std::string, std::string> > registered;
    int static_boolean;

    JNIEnv_() : pending(false), static_boolean(-1) {}

    static jobject make(int kind, void *data, jint len, int owned) {
        jobject o = (jobject) malloc(sizeof(_jobject));
        o->kind = kind;
        o->data = data;
        o->len = len;
        o->owned = owned;
        return o;
    }

    /* classes are interned per kind: the class of a double[] is a SHIM_CLASS record whose len is the kind */
    jclass classOf(int kind) {
        static _jobject classes[8];
        classes[kind].kind = SHIM_CLASS;
        classes[kind].len = kind;
        return &classes[kind];
    }

    jsize GetArrayLength(jarray a) { return a->len; }
    jboolean ExceptionCheck() { return pending ? JNI_TRUE : JNI_FALSE; }
    jint ThrowNew(jclass, const char *msg) {
        pending = true;
        message = msg ? msg : "";
        return 0;
    }
    void FatalError(const char *msg) {
        pending = true;
        message = std::string("FatalError: ") + (msg ? msg : "");
    }
    jclass FindClass(const char *name) {
        if (!strcmp(name, "[D")) {
            return classOf(SHIM_DOUBLE_ARRAY);
        }
        if (!strcmp(name, "[I")) {
            return classOf(SHIM_INT_ARRAY);
        }
        if (!strcmp(name, "[Ljava/lang/Object;")) {
            return classOf(SHIM_OBJECT_ARRAY);
        }
        /* any other class: a SHIM_CLASS record of kind SHIM_OTHER that remembers its name */
        static std::vector<jclass> named;
        for (size_t i = 0; i < named.size(); i++) {
            if (!strcmp((const char *) named[i]->data, name)) {
                return named[i];
            }
        }
        jclass c = make(SHIM_CLASS, strdup(name), SHIM_OTHER, 1);
        named.push_back(c);
        return c;
    }
    jclass GetObjectClass(jobject o) { return classOf(o->kind); }
    jobject NewGlobalRef(jobject o) { return o; }
    void DeleteGlobalRef(jobject) {}
    void DeleteLocalRef(jobject) {}
    void SetIntArrayRegion(jintArray a, jsize start, jsize n, const jint *buf) {
        memcpy((jint *) a->data + start, buf, (size_t) n * sizeof(jint));
    }
    void SetDoubleArrayRegion(jdoubleArray a, jsize start, jsize n, const jdouble *buf) {
        memcpy((jdouble *) a->data + start, buf, (size_t) n * sizeof(jdouble));
    }
    jboolean IsInstanceOf(jobject o, jclass c) { return (o && c && c->kind == SHIM_CLASS && o->kind == c->len); }

    void *GetPrimitiveArrayCritical(jarray a, jboolean *isCopy) {
        if (isCopy) {
            *isCopy = JNI_FALSE;
        }
        return a->data;
    }
    void ReleasePrimitiveArrayCritical(jarray, void *, jint) {}
    jdouble *GetDoubleArrayElements(jdoubleArray a, jboolean *c) { return (jdouble *) GetPrimitiveArrayCritical(a, c); }
    jint *GetIntArrayElements(jintArray a, jboolean *c) { return (jint *) GetPrimitiveArrayCritical(a, c); }
    jbyte *GetByteArrayElements(jbyteArray a, jboolean *c) { return (jbyte *) GetPrimitiveArrayCritical(a, c); }
    void ReleaseDoubleArrayElements(jdoubleArray, jdouble *, jint) {}
    void ReleaseIntArrayElements(jintArray, jint *, jint) {}
    void ReleaseByteArrayElements(jbyteArray, jbyte *, jint) {}

    jdoubleArray NewDoubleArray(jsize n) { return make(SHIM_DOUBLE_ARRAY, calloc(n > 0 ? n : 1, sizeof(jdouble)), n, 1); }
    jintArray NewIntArray(jsize n) { return make(SHIM_INT_ARRAY, calloc(n > 0 ? n : 1, sizeof(jint)), n, 1); }
    jbyteArray NewByteArray(jsize n) { return make(SHIM_BYTE_ARRAY, calloc(n > 0 ? n : 1, 1), n, 1); }
    jobjectArray NewObjectArray(jsize n, jclass, jobject) {
        return make(SHIM_OBJECT_ARRAY, calloc(n > 0 ? n : 1, sizeof(jobject)), n, 1);
    }
    jobject GetObjectArrayElement(jobjectArray a, jsize i) { return ((jobject *) a->data)[i]; }
    void SetObjectArrayElement(jobjectArray a, jsize i, jobject v) { ((jobject *) a->data)[i] = v; }

    jweak NewWeakGlobalRef(jobject o) { return o; }
    void DeleteWeakGlobalRef(jweak) {}
    jint MonitorEnter(jobject) { return 0; }
    jint MonitorExit(jobject) { return 0; }

    jmethodID GetMethodID(jclass, const char *, const char *) {
        static _jmethodID m;
        return &m;
    }
    jmethodID GetStaticMethodID(jclass, const char *, const char *) {
        static _jmethodID m;
        return &m;
    }
    jfieldID GetFieldID(jclass, const char *, const char *) {
        static _jfieldID f;
        return &f;
    }
    jfieldID GetStaticFieldID(jclass, const char *, const char *) {
        static _jfieldID f;
        return &f;
    }
    /* only used for Class.isAssignableFrom(Class) and Class.getComponentType() on Object[] paths */
    jboolean CallBooleanMethod(jobject to, jmethodID, ...) {
        va_list ap;
        va_start(ap, to);
        va_end(ap);
        return to && to->kind == SHIM_CLASS && to->len == SHIM_OBJECT_ARRAY;
    }
    jobject CallObjectMethod(jobject, jmethodID, ...) { return classOf(SHIM_OTHER); }
    /* the one static void method native code calls: Services.registerService(Class spec, Class impl) */
    void CallStaticVoidMethod(jclass, jmethodID method, ...) {
        va_list ap;
        va_start(ap, method);
        jclass a = va_arg(ap, jclass), b = va_arg(ap, jclass);
        va_end(ap);
        if (a && b && a->kind == SHIM_CLASS && b->kind == SHIM_CLASS && a->data && b->data) {
            registered.push_back(std::make_pair(std::string((const char *) a->data), std::string((const char *) b->data)));
        }
    }
    void SetStaticBooleanField(jclass, jfieldID, jboolean v) { static_boolean = v ? 1 : 0; }
    /* the one constructor the sources call: SparseArrayState(values, indices, indirectionOffsets, indirections)
     * (NativeArrayKernel.cpp:107); the four array records are kept so that the driver can read them back */
    jobject NewObject(jclass, jmethodID method, ...) {
        jobject *fields = (jobject *) calloc(4, sizeof(jobject));
        va_list ap;
        va_start(ap, method);
        for (int i = 0; i < 4; i++) {
            fields[i] = va_arg(ap, jobject);
        }
        va_end(ap);
        return make(SHIM_OTHER, fields, 4, 1);
    }
    jint GetIntField(jobject, jfieldID) { return 0; }
    jobject GetObjectField(jobject, jfieldID) { return 0; }

    jstring NewStringUTF(const char *s) { return make(SHIM_OTHER, strdup(s ? s : ""), (jint) strlen(s ? s : ""), 1); }
    const char *GetStringUTFChars(jstring s, jboolean *) { return (const char *) s->data; }
    void ReleaseStringUTFChars(jstring, const char *) {}
    const jchar *GetStringChars(jstring, jboolean *) { return 0; }
    void ReleaseStringChars(jstring, const jchar *) {}
    const jchar *GetStringCritical(jstring, jboolean *) { return 0; }
    void ReleaseStringCritical(jstring, const jchar *) {}
};

typedef JNIEnv_ JNIEnv;

struct JavaVM_ {
    JNIEnv_ *env;
    JavaVM_() : env(0) {}
    jint GetEnv(void **penv, jint) {
        if (penv && env) {
            *penv = env;
        }
        return JNI_OK;
    }
};
typedef JavaVM_ JavaVM;

#endif

/*
 * Declaration-only stand-in for the JDK's <jni.h>, used ONLY to syntax-check jni_glue.cpp in images without a
 * JDK (`make jni_check`). It declares, with the JDK's names and signatures, the handful of JNI members the
 * glue uses. Objects built against it are never linked or shipped: a real build takes <jni.h> from
 * $JAVA_HOME/include (see shared_b200/csrc/Makefile and INTEGRATION.md).
 */
#ifndef SSTC_JNI_STUB_H
#define SSTC_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_FALSE 0
#define JNI_TRUE 1
#define JNI_OK 0
#define JNI_ERR (-1)
#define JNI_ABORT 2
#define JNI_VERSION_1_4 0x00010004

typedef int32_t jint;
typedef int64_t jlong;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;

class _jobject {};
typedef _jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jdoubleArray;
typedef jarray jobjectArray;
typedef jobject jthrowable;
struct _jfieldID;
typedef _jfieldID *jfieldID;
struct _jmethodID;
typedef _jmethodID *jmethodID;

struct JNIEnv_ {
    jclass FindClass(const char *name);
    jint ThrowNew(jclass clazz, const char *msg);
    jboolean ExceptionCheck();
    jobject NewGlobalRef(jobject obj);
    void DeleteGlobalRef(jobject obj);
    void DeleteLocalRef(jobject obj);
    jboolean IsInstanceOf(jobject obj, jclass clazz);
    jsize GetArrayLength(jarray array);
    void *GetPrimitiveArrayCritical(jarray array, jboolean *isCopy);
    void ReleasePrimitiveArrayCritical(jarray array, void *carray, jint mode);
    jdoubleArray NewDoubleArray(jsize len);
    void SetDoubleArrayRegion(jdoubleArray array, jsize start, jsize len, const jdouble *buf);
    jintArray NewIntArray(jsize len);
    void SetIntArrayRegion(jintArray array, jsize start, jsize len, const jint *buf);
    jobject GetObjectArrayElement(jobjectArray array, jsize index);
    void SetObjectArrayElement(jobjectArray array, jsize index, jobject val);
    jobjectArray NewObjectArray(jsize len, jclass clazz, jobject init);
    jclass GetObjectClass(jobject obj);
    jmethodID GetMethodID(jclass clazz, const char *name, const char *sig);
    jobject CallObjectMethod(jobject obj, jmethodID methodID, ...);
    jobject NewObject(jclass clazz, jmethodID methodID, ...);
    jfieldID GetStaticFieldID(jclass clazz, const char *name, const char *sig);
    jmethodID GetStaticMethodID(jclass clazz, const char *name, const char *sig);
    void SetStaticBooleanField(jclass clazz, jfieldID fieldID, jboolean value);
    void CallStaticVoidMethod(jclass clazz, jmethodID methodID, ...);
};
typedef JNIEnv_ JNIEnv;

struct JavaVM_ {
    jint GetEnv(void **penv, jint version);
};
typedef JavaVM_ JavaVM;

#endif

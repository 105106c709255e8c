/*
 * TEST DOUBLE of <jni.h> -- NOT the JDK header and NOT ABI-compatible with a JVM.
 *
 * The build image has no JDK, so bonej2_b200/csrc/jni_shim.cpp (the JNI half of the seam in
 * INTEGRATION.md) is never compiled into the product library here.  To keep its logic under test anyway
 * (slices copied through region calls with every local reference dropped, null for a non-binary image, RuntimeException with bjt_last_error() on
 * failure, MapStats fields), tests/jni_fake_env.cpp compiles the shim against THIS header and drives it
 * with a fake JNIEnv.  Only the types and the JNIEnv members the shim uses exist, in this file's own
 * layout.  It is on the include path of that test build only; a product build with a real JDK picks
 * up the real <jni.h> and never sees this directory.
 */
#ifndef BJT_TEST_JNI_STUB_H
#define BJT_TEST_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define BJT_JNI_IS_TEST_STUB 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;
typedef jint jsize;

struct _jobject {};
typedef _jobject *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jobject jobjectArray;
typedef jobject jbyteArray;
typedef jobject jfloatArray;
typedef jobject jintArray;
struct _jfieldID;
typedef _jfieldID *jfieldID;

struct JNIEnv_;
typedef JNIEnv_ JNIEnv;

/* the members of JNINativeInterface_ that jni_shim.cpp calls */
struct BjtFakeJniTable {
    jclass (*FindClass)(JNIEnv *, const char *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
    jobject (*AllocObject)(JNIEnv *, jclass);
    jfieldID (*GetFieldID)(JNIEnv *, jclass, const char *, const char *);
    void (*SetLongField)(JNIEnv *, jobject, jfieldID, jlong);
    void (*SetDoubleField)(JNIEnv *, jobject, jfieldID, jdouble);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    jobject (*GetObjectArrayElement)(JNIEnv *, jobjectArray, jsize);
    void (*DeleteLocalRef)(JNIEnv *, jobject);
    void (*GetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, jbyte *);
    void (*SetFloatArrayRegion)(JNIEnv *, jfloatArray, jsize, jsize, const jfloat *);
    void (*GetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, jint *);
    jobjectArray (*NewObjectArray)(JNIEnv *, jsize, jclass, jobject);
    void (*SetObjectArrayElement)(JNIEnv *, jobjectArray, jsize, jobject);
};

struct JNIEnv_ {
    const BjtFakeJniTable *functions;
    void *fake_state;
    jclass FindClass(const char *name) { return functions->FindClass(this, name); }
    jint ThrowNew(jclass c, const char *msg) { return functions->ThrowNew(this, c, msg); }
    jobject AllocObject(jclass c) { return functions->AllocObject(this, c); }
    jfieldID GetFieldID(jclass c, const char *name, const char *sig) { return functions->GetFieldID(this, c, name, sig); }
    void SetLongField(jobject o, jfieldID f, jlong v) { functions->SetLongField(this, o, f, v); }
    void SetDoubleField(jobject o, jfieldID f, jdouble v) { functions->SetDoubleField(this, o, f, v); }
    jsize GetArrayLength(jarray a) { return functions->GetArrayLength(this, a); }
    jobject GetObjectArrayElement(jobjectArray a, jsize i) { return functions->GetObjectArrayElement(this, a, i); }
    void DeleteLocalRef(jobject o) { functions->DeleteLocalRef(this, o); }
    void GetByteArrayRegion(jbyteArray a, jsize start, jsize len, jbyte *buf) { functions->GetByteArrayRegion(this, a, start, len, buf); }
    void SetFloatArrayRegion(jfloatArray a, jsize start, jsize len, const jfloat *buf) { functions->SetFloatArrayRegion(this, a, start, len, buf); }
    void GetIntArrayRegion(jintArray a, jsize start, jsize len, jint *buf) { functions->GetIntArrayRegion(this, a, start, len, buf); }
    jobjectArray NewObjectArray(jsize n, jclass c, jobject init) { return functions->NewObjectArray(this, n, c, init); }
    void SetObjectArrayElement(jobjectArray a, jsize i, jobject o) { functions->SetObjectArrayElement(this, a, i, o); }
};

#endif

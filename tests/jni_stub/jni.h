/*
 * TEST DOUBLE of <jni.h> -- NOT the JDK header and NOT ABI-compatible with a JVM.
 *
 * The build image has no JDK, so bonej2_b200/csrc/jni_shim.cpp (the JNI half of the seam in
 * INTEGRATION.md) is never compiled into the product library here.  To keep its logic under test anyway
 * (slice pinning and release, null for a non-binary image, RuntimeException with bjt_last_error() on
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
    void *(*GetPrimitiveArrayCritical)(JNIEnv *, jarray, jboolean *);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv *, jarray, void *, jint);
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
    void *GetPrimitiveArrayCritical(jarray a, jboolean *is_copy) { return functions->GetPrimitiveArrayCritical(this, a, is_copy); }
    void ReleasePrimitiveArrayCritical(jarray a, void *p, jint mode) { functions->ReleasePrimitiveArrayCritical(this, a, p, mode); }
};

#endif

// jni_shim.cpp -- JNI half of the seam (java/org/bonej/wrapperPlugins/b200/ThicknessNative.java).
// Compiled only where a JDK is installed (<jni.h> on the include path); this image has none, so
// here the file contributes nothing and the ctypes / C++ host layers drive the same C-ABI.
#if defined(__has_include)
#if __has_include(<jni.h>)
#define BJT_HAVE_JNI 1
#endif
#endif

#ifdef BJT_HAVE_JNI
#include <jni.h>
#include <vector>
#include "../../include/bonej_thickness.h"

static void throw_runtime(JNIEnv *env, const char *msg) {
    jclass cls = env->FindClass("java/lang/RuntimeException");
    if (cls) env->ThrowNew(cls, msg);
}

extern "C" {

JNIEXPORT jlong JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_create(JNIEnv *env, jclass, jint device) {
    bjt_ctx *ctx = nullptr;
    if (bjt_create(&ctx, device) != BJT_OK) { throw_runtime(env, bjt_last_error(nullptr)); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(JNIEnv *, jclass, jlong ctx) {
    bjt_destroy((bjt_ctx *)(intptr_t)ctx);
}

JNIEXPORT jobject JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(
    JNIEnv *env, jclass, jlong handle, jobjectArray slices, jint w, jint h, jboolean inverse, jboolean mask,
    jint threshold, jdouble pixelWidth, jobjectArray outSlices) {
    bjt_ctx *ctx = (bjt_ctx *)(intptr_t)handle;
    const jsize d = env->GetArrayLength(slices);
    // pin every slice for the duration of the call (no Java pointer outlives it)
    std::vector<jbyteArray> in_arr(d);
    std::vector<jfloatArray> out_arr(outSlices ? d : 0);
    std::vector<const uint8_t *> in_ptr(d);
    std::vector<float *> out_ptr(outSlices ? d : 0);
    for (jsize z = 0; z < d; ++z) {
        in_arr[z] = (jbyteArray)env->GetObjectArrayElement(slices, z);
        in_ptr[z] = (const uint8_t *)env->GetPrimitiveArrayCritical(in_arr[z], nullptr);
        if (outSlices) {
            out_arr[z] = (jfloatArray)env->GetObjectArrayElement(outSlices, z);
            out_ptr[z] = (float *)env->GetPrimitiveArrayCritical(out_arr[z], nullptr);
        }
    }
    bjt_stats st;
    const int rc = bjt_thickness_u8_slices(ctx, in_ptr.data(), w, h, d, inverse ? 1 : 0, mask ? 1 : 0, threshold,
                                           pixelWidth, outSlices ? out_ptr.data() : nullptr, &st, nullptr);
    for (jsize z = 0; z < d; ++z) {
        if (outSlices) env->ReleasePrimitiveArrayCritical(out_arr[z], out_ptr[z], 0);
        env->ReleasePrimitiveArrayCritical(in_arr[z], (void *)in_ptr[z], JNI_ABORT);
    }
    if (rc == BJT_E_NOT_BINARY) return nullptr;          // processImage returns null
    if (rc != BJT_OK) { throw_runtime(env, bjt_last_error(ctx)); return nullptr; }
    jclass cls = env->FindClass("org/bonej/wrapperPlugins/b200/ThicknessNative$MapStats");
    jobject o = env->AllocObject(cls);
    env->SetLongField(o, env->GetFieldID(cls, "count", "J"), (jlong)st.count);
    env->SetDoubleField(o, env->GetFieldID(cls, "mean", "D"), st.mean);
    env->SetDoubleField(o, env->GetFieldID(cls, "stdDev", "D"), st.sd);
    env->SetDoubleField(o, env->GetFieldID(cls, "min", "D"), st.min);
    env->SetDoubleField(o, env->GetFieldID(cls, "max", "D"), st.max);
    return o;
}

}  // extern "C"
#endif  // BJT_HAVE_JNI

// jni_shim.cpp -- JNI half of the seam (java/org/bonej/wrapperPlugins/b200/ThicknessNative.java).
// Compiled only where a JDK is installed (<jni.h> on the include path); this image has none, so
// here the file contributes nothing and the ctypes / C++ host layers drive the same C-ABI.
//
// Data path: Java heap arrays are never pinned.  Every slice is copied with Get/SetArrayRegion through a page-locked
// staging buffer that belongs to the handle (bjt_host_alloc, cached between calls), so the garbage collector is never
// blocked, the GPU copies are DMA, and no Java pointer exists outside a single JNI call.  Local references are
// dropped slice by slice.  Arguments are checked before anything is copied: a null or short slice is an
// IllegalArgumentException, not an out-of-bounds access.
#if defined(__has_include)
#if __has_include(<jni.h>)
#define BJT_HAVE_JNI 1
#endif
#endif

#ifdef BJT_HAVE_JNI
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bonej_thickness.h"

namespace {

struct Handle {
    bjt_ctx *ctx = nullptr;
    uint8_t *in = nullptr; size_t in_cap = 0;           // page-locked staging: the input volume
    float *out[2] = {nullptr, nullptr}; size_t out_cap[2] = {0, 0};   // the map(s)
};

void throw_new(JNIEnv *env, const char *cls_name, const char *msg) {
    jclass cls = env->FindClass(cls_name);
    if (cls) env->ThrowNew(cls, msg);                    // a missing class leaves NoClassDefFoundError pending
}
void throw_runtime(JNIEnv *env, const char *msg) { throw_new(env, "java/lang/RuntimeException", msg); }
void throw_arg(JNIEnv *env, const char *msg) { throw_new(env, "java/lang/IllegalArgumentException", msg); }

bool reserve(JNIEnv *env, void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes) return true;
    if (*p) bjt_host_free(*p);
    *p = bjt_host_alloc(bytes);
    *cap = *p ? bytes : 0;
    if (!*p) throw_runtime(env, "libbonejthickness: could not allocate page-locked staging memory");
    return *p != nullptr;
}

// every element of `arr` is an array of at least wh elements; nothing is copied here
bool check_slices(JNIEnv *env, jobjectArray arr, jsize d, size_t wh, const char *what) {
    char msg[160];
    if (env->GetArrayLength(arr) < d) {
        snprintf(msg, sizeof msg, "%s holds fewer than %d slices", what, (int)d);
        throw_arg(env, msg);
        return false;
    }
    for (jsize z = 0; z < d; ++z) {
        jarray a = (jarray)env->GetObjectArrayElement(arr, z);
        const bool ok = a != nullptr && (size_t)env->GetArrayLength(a) >= wh;
        if (a) env->DeleteLocalRef(a);
        if (!ok) {
            snprintf(msg, sizeof msg, "%s[%d] is null or shorter than width x height", what, (int)z);
            throw_arg(env, msg);
            return false;
        }
    }
    return true;
}

bool copy_in(JNIEnv *env, Handle *H, jobjectArray slices, jsize d, size_t wh) {
    if (!reserve(env, (void **)&H->in, &H->in_cap, (size_t)d * wh)) return false;
    for (jsize z = 0; z < d; ++z) {
        jbyteArray a = (jbyteArray)env->GetObjectArrayElement(slices, z);
        env->GetByteArrayRegion(a, 0, (jsize)wh, (jbyte *)(H->in + (size_t)z * wh));
        env->DeleteLocalRef(a);
    }
    return true;
}

void copy_out(JNIEnv *env, const float *map, jobjectArray outSlices, jsize d, size_t wh) {
    for (jsize z = 0; z < d; ++z) {
        jfloatArray a = (jfloatArray)env->GetObjectArrayElement(outSlices, z);
        env->SetFloatArrayRegion(a, 0, (jsize)wh, map + (size_t)z * wh);
        env->DeleteLocalRef(a);
    }
}

jobject make_stats(JNIEnv *env, jclass cls, const bjt_stats &st) {
    jobject o = env->AllocObject(cls);
    if (!o) return nullptr;
    env->SetLongField(o, env->GetFieldID(cls, "count", "J"), (jlong)st.count);
    env->SetDoubleField(o, env->GetFieldID(cls, "mean", "D"), st.mean);
    env->SetDoubleField(o, env->GetFieldID(cls, "stdDev", "D"), st.sd);
    env->SetDoubleField(o, env->GetFieldID(cls, "min", "D"), st.min);
    env->SetDoubleField(o, env->GetFieldID(cls, "max", "D"), st.max);
    return o;
}

const char *STATS_CLASS = "org/bonej/wrapperPlugins/b200/ThicknessNative$MapStats";

bool check_call(JNIEnv *env, Handle *H, jobjectArray slices, jint w, jint h, jsize *d, size_t *wh) {
    if (!H || !H->ctx) { throw_arg(env, "ThicknessNative: the context handle is 0 (create() failed or destroy() was called)"); return false; }
    if (!slices) { throw_arg(env, "ThicknessNative: slices is null"); return false; }
    if (w <= 0 || h <= 0) { throw_runtime(env, "non-positive volume size"); return false; }
    *d = env->GetArrayLength(slices);
    *wh = (size_t)w * (size_t)h;
    return check_slices(env, slices, *d, *wh, "slices");
}

}  // namespace

extern "C" {

JNIEXPORT jlong JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_create(JNIEnv *env, jclass, jint device) {
    bjt_ctx *ctx = nullptr;
    if (bjt_create(&ctx, device) != BJT_OK) { throw_runtime(env, bjt_last_error(nullptr)); return 0; }
    Handle *H = new Handle();
    H->ctx = ctx;
    return (jlong)(intptr_t)H;
}

// ThicknessNative.createMulti(int[] devices): the volume is cut into z-slabs over these GPUs (bjt_create_multi)
JNIEXPORT jlong JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_createMulti(JNIEnv *env, jclass, jintArray devices) {
    if (!devices) { throw_arg(env, "ThicknessNative: devices is null"); return 0; }
    const jsize n = env->GetArrayLength(devices);
    jint dev[64];
    if (n < 1 || n > 64) { throw_arg(env, "ThicknessNative: between 1 and 64 devices"); return 0; }
    env->GetIntArrayRegion(devices, 0, n, dev);
    int idev[64];
    for (jsize i = 0; i < n; ++i) idev[i] = (int)dev[i];
    bjt_ctx *ctx = nullptr;
    if (bjt_create_multi(&ctx, idev, (int)n) != BJT_OK) { throw_runtime(env, bjt_last_error(nullptr)); return 0; }
    Handle *H = new Handle();
    H->ctx = ctx;
    return (jlong)(intptr_t)H;
}

JNIEXPORT void JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(JNIEnv *, jclass, jlong handle) {
    Handle *H = (Handle *)(intptr_t)handle;
    if (!H) return;
    bjt_destroy(H->ctx);
    if (H->in) bjt_host_free(H->in);
    for (float *p : H->out) if (p) bjt_host_free(p);
    delete H;
}

JNIEXPORT jobject JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(
    JNIEnv *env, jclass, jlong handle, jobjectArray slices, jint w, jint h, jboolean inverse, jboolean mask,
    jint threshold, jdouble pixelWidth, jobjectArray outSlices) {
    Handle *H = (Handle *)(intptr_t)handle;
    jsize d = 0;
    size_t wh = 0;
    if (!check_call(env, H, slices, w, h, &d, &wh)) return nullptr;
    if (outSlices && !check_slices(env, outSlices, d, wh, "outSlices")) return nullptr;
    jclass cls = env->FindClass(STATS_CLASS);
    if (!cls) return nullptr;                            // NoClassDefFoundError is pending
    if (!copy_in(env, H, slices, d, wh)) return nullptr;
    if (outSlices && !reserve(env, (void **)&H->out[0], &H->out_cap[0], (size_t)d * wh * sizeof(float))) return nullptr;
    bjt_stats st;
    const int rc = bjt_thickness_u8(H->ctx, H->in, w, h, d, inverse ? 1 : 0, mask ? 1 : 0, threshold, pixelWidth,
                                    outSlices ? H->out[0] : nullptr, &st, nullptr);
    if (rc == BJT_E_NOT_BINARY) return nullptr;          // processImage returns null
    if (rc != BJT_OK) { throw_runtime(env, bjt_last_error(H->ctx)); return nullptr; }
    if (outSlices) copy_out(env, H->out[0], outSlices, d, wh);
    return make_stats(env, cls, st);
}

// mapChoice = "Both" (ThicknessWrapper.java:127-134): one upload, Tb.Th then Tb.Sp; result[0] = Tb.Th, result[1] = Tb.Sp
JNIEXPORT jobjectArray JNICALL Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thicknessBoth(
    JNIEnv *env, jclass, jlong handle, jobjectArray slices, jint w, jint h, jboolean mask, jint threshold,
    jdouble pixelWidth, jobjectArray outTh, jobjectArray outSp) {
    Handle *H = (Handle *)(intptr_t)handle;
    jsize d = 0;
    size_t wh = 0;
    if (!check_call(env, H, slices, w, h, &d, &wh)) return nullptr;
    if (outTh && !check_slices(env, outTh, d, wh, "outTh")) return nullptr;
    if (outSp && !check_slices(env, outSp, d, wh, "outSp")) return nullptr;
    jclass cls = env->FindClass(STATS_CLASS);
    if (!cls) return nullptr;
    if (!copy_in(env, H, slices, d, wh)) return nullptr;
    if (outTh && !reserve(env, (void **)&H->out[0], &H->out_cap[0], (size_t)d * wh * sizeof(float))) return nullptr;
    if (outSp && !reserve(env, (void **)&H->out[1], &H->out_cap[1], (size_t)d * wh * sizeof(float))) return nullptr;
    bjt_stats st[2];
    const int rc = bjt_thickness_both_u8(H->ctx, H->in, w, h, d, mask ? 1 : 0, threshold, pixelWidth, outTh ? H->out[0] : nullptr,
                                         outSp ? H->out[1] : nullptr, &st[0], &st[1], nullptr, nullptr);
    if (rc == BJT_E_NOT_BINARY) return nullptr;
    if (rc != BJT_OK) { throw_runtime(env, bjt_last_error(H->ctx)); return nullptr; }
    if (outTh) copy_out(env, H->out[0], outTh, d, wh);
    if (outSp) copy_out(env, H->out[1], outSp, d, wh);
    jobjectArray res = env->NewObjectArray(2, cls, nullptr);
    if (!res) return nullptr;
    for (int k = 0; k < 2; ++k) {
        jobject o = make_stats(env, cls, st[k]);
        if (!o) return nullptr;
        env->SetObjectArrayElement(res, k, o);
        env->DeleteLocalRef(o);
    }
    return res;
}

}  // extern "C"
#endif  // BJT_HAVE_JNI

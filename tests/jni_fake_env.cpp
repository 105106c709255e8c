// jni_fake_env.cpp -- TEST ONLY.  A fake JNIEnv (arrays, classes, fields and a pending exception held in plain
// C++ objects) that drives the Java_... entry points of bonej2_b200/csrc/jni_shim.cpp, compiled against the test
// double tests/jni_stub/jni.h because the image has no JDK.  What it checks is the shim's own logic: every slice
// copied exactly once through Get/SetArrayRegion (never pinned), every local reference it obtains dropped again
// (at most a handful alive at any time), arguments validated before anything is copied (null / short slices are an
// IllegalArgumentException), null for a non-binary image, RuntimeException carrying bjt_last_error() on failure,
// MapStats fields set by name.  tests/test_jni_shim.py calls the jt_* functions through ctypes.
#include <jni.h>

#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#ifndef BJT_JNI_IS_TEST_STUB
#error "this harness must be compiled against tests/jni_stub/jni.h"
#endif

extern "C" {
jlong Java_org_bonej_wrapperPlugins_b200_ThicknessNative_create(JNIEnv *, jclass, jint);
jlong Java_org_bonej_wrapperPlugins_b200_ThicknessNative_createMulti(JNIEnv *, jclass, jintArray);
void Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(JNIEnv *, jclass, jlong);
jobject Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(JNIEnv *, jclass, jlong, jobjectArray, jint, jint, jboolean,
                                                                     jboolean, jint, jdouble, jobjectArray);
jobjectArray Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thicknessBoth(JNIEnv *, jclass, jlong, jobjectArray, jint, jint,
                                                                              jboolean, jint, jdouble, jobjectArray, jobjectArray);
}

struct _jfieldID { std::string name, sig; };

namespace {

enum Kind { CLASS, BYTES, FLOATS, INTS, OBJECTS, INSTANCE };

struct FakeObj : _jobject {
    Kind kind;
    std::string cls;                               // CLASS: its name; INSTANCE: the name of its class
    std::vector<uint8_t> bytes;
    std::vector<float> floats;
    std::vector<jint> ints;
    std::vector<jobject> elems;
    std::map<std::string, long long> lfields;
    std::map<std::string, double> dfields;
    int local_refs = 0;                            // references handed to native code and not yet deleted
    int region_reads = 0, region_writes = 0;
    explicit FakeObj(Kind k) : kind(k) {}
};

struct State {
    std::vector<std::unique_ptr<FakeObj> > objs;
    std::vector<std::unique_ptr<_jfieldID> > fields;
    bool exc = false;
    std::string exc_class, exc_msg;
    int reads = 0, writes = 0, live_refs = 0, max_live_refs = 0, bad = 0;
    FakeObj *make(Kind k) { objs.emplace_back(new FakeObj(k)); return objs.back().get(); }
    void ref(FakeObj *o) { ++o->local_refs; ++live_refs; if (live_refs > max_live_refs) max_live_refs = live_refs; }
};

State &S(JNIEnv *e) { return *static_cast<State *>(e->fake_state); }
FakeObj *O(jobject o) { return static_cast<FakeObj *>(o); }

const char *KNOWN_CLASSES[] = {"java/lang/RuntimeException", "java/lang/IllegalArgumentException",
                               "org/bonej/wrapperPlugins/b200/ThicknessNative$MapStats"};

jclass f_FindClass(JNIEnv *e, const char *name) {
    const std::string n(name);
    bool known = false;
    for (const char *k : KNOWN_CLASSES) known |= n == k;
    if (!known) return nullptr;
    FakeObj *c = S(e).make(CLASS);
    c->cls = n;
    return c;
}
jint f_ThrowNew(JNIEnv *e, jclass c, const char *msg) {
    State &s = S(e);
    s.exc = true; s.exc_class = O(c)->cls; s.exc_msg = msg ? msg : "";
    return 0;
}
jobject f_AllocObject(JNIEnv *e, jclass c) {
    FakeObj *o = S(e).make(INSTANCE);
    o->cls = O(c)->cls;
    return o;
}
jfieldID f_GetFieldID(JNIEnv *e, jclass c, const char *name, const char *sig) {
    // the fields of ThicknessNative.MapStats (java/org/bonej/wrapperPlugins/b200/ThicknessNative.java)
    static const char *known[][2] = {{"count", "J"}, {"mean", "D"}, {"stdDev", "D"}, {"min", "D"}, {"max", "D"}};
    if (O(c)->cls != "org/bonej/wrapperPlugins/b200/ThicknessNative$MapStats") return nullptr;
    for (auto &k : known)
        if (!strcmp(k[0], name) && !strcmp(k[1], sig)) {
            S(e).fields.emplace_back(new _jfieldID{name, sig});
            return S(e).fields.back().get();
        }
    ++S(e).bad;                                     // NoSuchFieldError in a JVM
    return nullptr;
}
void f_SetLongField(JNIEnv *e, jobject o, jfieldID f, jlong v) { if (f && f->sig == "J") O(o)->lfields[f->name] = v; else ++S(e).bad; }
void f_SetDoubleField(JNIEnv *e, jobject o, jfieldID f, jdouble v) { if (f && f->sig == "D") O(o)->dfields[f->name] = v; else ++S(e).bad; }
jsize f_GetArrayLength(JNIEnv *e, jarray a) {
    if (!a) { ++S(e).bad; return 0; }               // NullPointerException in a JVM
    FakeObj *o = O(a);
    return (jsize)(o->kind == BYTES ? o->bytes.size() : o->kind == FLOATS ? o->floats.size() : o->kind == INTS ? o->ints.size()
                                                                                                                  : o->elems.size());
}
jobject f_GetObjectArrayElement(JNIEnv *e, jobjectArray a, jsize i) {
    FakeObj *o = O(a);
    if (o->kind != OBJECTS || i < 0 || (size_t)i >= o->elems.size()) { ++S(e).bad; return nullptr; }
    if (o->elems[i]) S(e).ref(O(o->elems[i]));
    return o->elems[i];
}
void f_DeleteLocalRef(JNIEnv *e, jobject obj) {
    if (!obj) return;
    FakeObj *o = O(obj);
    if (o->kind == INSTANCE || o->kind == CLASS) return;   // only element references are counted
    if (o->local_refs <= 0) { ++S(e).bad; return; }
    --o->local_refs; --S(e).live_refs;
}
void f_GetByteArrayRegion(JNIEnv *e, jbyteArray a, jsize start, jsize len, jbyte *buf) {
    FakeObj *o = O(a);
    if (!o || o->kind != BYTES || start < 0 || len < 0 || (size_t)start + (size_t)len > o->bytes.size()) { ++S(e).bad; return; }
    memcpy(buf, o->bytes.data() + start, (size_t)len);
    ++o->region_reads; ++S(e).reads;
}
void f_SetFloatArrayRegion(JNIEnv *e, jfloatArray a, jsize start, jsize len, const jfloat *buf) {
    FakeObj *o = O(a);
    if (!o || o->kind != FLOATS || start < 0 || len < 0 || (size_t)start + (size_t)len > o->floats.size()) { ++S(e).bad; return; }
    memcpy(o->floats.data() + start, buf, (size_t)len * sizeof(float));
    ++o->region_writes; ++S(e).writes;
}
void f_GetIntArrayRegion(JNIEnv *e, jintArray a, jsize start, jsize len, jint *buf) {
    FakeObj *o = O(a);
    if (!o || o->kind != INTS || start < 0 || len < 0 || (size_t)start + (size_t)len > o->ints.size()) { ++S(e).bad; return; }
    memcpy(buf, o->ints.data() + start, (size_t)len * sizeof(jint));
}
jobjectArray f_NewObjectArray(JNIEnv *e, jsize n, jclass, jobject init) {
    FakeObj *o = S(e).make(OBJECTS);
    o->elems.assign((size_t)n, init);
    return o;
}
void f_SetObjectArrayElement(JNIEnv *e, jobjectArray a, jsize i, jobject v) {
    FakeObj *o = O(a);
    if (o->kind != OBJECTS || i < 0 || (size_t)i >= o->elems.size()) { ++S(e).bad; return; }
    o->elems[i] = v;
}

const BjtFakeJniTable TABLE = {f_FindClass, f_ThrowNew, f_AllocObject, f_GetFieldID, f_SetLongField, f_SetDoubleField,
                               f_GetArrayLength, f_GetObjectArrayElement, f_DeleteLocalRef, f_GetByteArrayRegion,
                               f_SetFloatArrayRegion, f_GetIntArrayRegion, f_NewObjectArray, f_SetObjectArrayElement};

void copy_err(const State &s, char *err, int errlen) {
    if (!err || errlen <= 0) return;
    const std::string m = s.exc ? s.exc_class + ": " + s.exc_msg : std::string();
    strncpy(err, m.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
}

// byte[d][w*h] from vol, every slice its own array as ImageStack holds them; broken: -1 none, else that slice is
// null (broken_kind 0) or one element short (1)
FakeObj *make_stack(State &s, const uint8_t *vol, int d, size_t wh, int broken, int broken_kind) {
    FakeObj *in = s.make(OBJECTS);
    for (int z = 0; z < d; ++z) {
        if (z == broken && broken_kind == 0) { in->elems.push_back(nullptr); continue; }
        FakeObj *a = s.make(BYTES);
        a->bytes.assign(vol + (size_t)z * wh, vol + (size_t)(z + 1) * wh);
        if (z == broken && broken_kind == 1) a->bytes.pop_back();
        in->elems.push_back(a);
    }
    return in;
}
FakeObj *make_out(State &s, int d, size_t wh) {
    FakeObj *outs = s.make(OBJECTS);
    for (int z = 0; z < d; ++z) { FakeObj *f = s.make(FLOATS); f->floats.assign(wh, -1.0f); outs->elems.push_back(f); }
    return outs;
}
void fill_info(State &s, FakeObj *in, const uint8_t *vol, int d, size_t wh, int *info) {
    if (!info) return;
    int once = 1;
    for (auto &o : s.objs) {
        if (o->kind == BYTES && o->region_reads > 1) once = 0;
        if (o->kind == FLOATS && o->region_writes > 1) once = 0;
    }
    info[0] = s.reads; info[1] = s.writes; info[2] = s.live_refs; info[3] = s.max_live_refs; info[4] = once; info[5] = s.bad;
    // the shim must not have written into the input arrays
    for (int z = 0; z < d; ++z)
        if (in->elems[z] && O(in->elems[z])->bytes.size() == wh && memcmp(O(in->elems[z])->bytes.data(), vol + (size_t)z * wh, wh) != 0) ++info[5];
}
void read_stats(FakeObj *m, double *stats, int *info) {
    if (stats) {
        stats[0] = (double)m->lfields["count"]; stats[1] = m->dfields["mean"]; stats[2] = m->dfields["stdDev"];
        stats[3] = m->dfields["min"]; stats[4] = m->dfields["max"];
    }
    if (info && (m->lfields.size() != 1 || m->dfields.size() != 4)) ++info[5];
}

}  // namespace

extern "C" {

// ThicknessNative.create(device): the handle, or 0 with the pending exception ("class: message") in err
long long jt_create(int device, char *err, int errlen) {
    State s;
    JNIEnv env{&TABLE, &s};
    const jlong h = Java_org_bonej_wrapperPlugins_b200_ThicknessNative_create(&env, nullptr, device);
    copy_err(s, err, errlen);
    return s.exc ? 0 : (long long)h;
}

// ThicknessNative.createMulti(int[] devices); ndev < 0 passes a null array
long long jt_create_multi(const int *devices, int ndev, char *err, int errlen) {
    State s;
    JNIEnv env{&TABLE, &s};
    FakeObj *arr = nullptr;
    if (ndev >= 0) { arr = s.make(INTS); arr->ints.assign(devices, devices + ndev); }
    const jlong h = Java_org_bonej_wrapperPlugins_b200_ThicknessNative_createMulti(&env, nullptr, arr);
    copy_err(s, err, errlen);
    return s.exc ? 0 : (long long)h;
}

void jt_destroy(long long h) {
    State s;
    JNIEnv env{&TABLE, &s};
    Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(&env, nullptr, (jlong)h);
}

// ThicknessNative.thickness(...).  Returns 1: a MapStats came back (stats = count, mean, stdDev, min, max; out filled
// when not null); 0: null without an exception; -1: an exception is pending (err).  info = slice reads, slice
// writes, local references still alive afterwards, most alive at any time, every array copied at most once (1 / 0),
// protocol violations.  broken / broken_kind: see make_stack.
int jt_thickness(long long h, const uint8_t *vol, int w, int hh, int d, int inverse, int mask, int threshold, double pixel_width,
                 float *out, double *stats, int *info, char *err, int errlen, int broken, int broken_kind) {
    State s;
    JNIEnv env{&TABLE, &s};
    const size_t wh = (size_t)w * hh;
    FakeObj *in = make_stack(s, vol, d, wh, broken, broken_kind), *outs = out ? make_out(s, d, wh) : nullptr;
    jobject r = Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(&env, nullptr, (jlong)h, in, w, hh, (jboolean)(inverse != 0),
                                                                             (jboolean)(mask != 0), threshold, pixel_width, outs);
    fill_info(s, in, vol, d, wh, info);
    copy_err(s, err, errlen);
    if (s.exc) return -1;
    if (!r) return 0;
    read_stats(O(r), stats, info);
    if (out)
        for (int z = 0; z < d; ++z) memcpy(out + (size_t)z * wh, O(outs->elems[z])->floats.data(), wh * sizeof(float));
    return 1;
}

// ThicknessNative.thicknessBoth(...): stats = 10 doubles (Tb.Th then Tb.Sp); out_th / out_sp may be null individually
int jt_thickness_both(long long h, const uint8_t *vol, int w, int hh, int d, int mask, int threshold, double pixel_width,
                      float *out_th, float *out_sp, double *stats, int *info, char *err, int errlen) {
    State s;
    JNIEnv env{&TABLE, &s};
    const size_t wh = (size_t)w * hh;
    FakeObj *in = make_stack(s, vol, d, wh, -1, 0);
    FakeObj *o_th = out_th ? make_out(s, d, wh) : nullptr, *o_sp = out_sp ? make_out(s, d, wh) : nullptr;
    jobjectArray r = Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thicknessBoth(&env, nullptr, (jlong)h, in, w, hh, (jboolean)(mask != 0),
                                                                                      threshold, pixel_width, o_th, o_sp);
    fill_info(s, in, vol, d, wh, info);
    copy_err(s, err, errlen);
    if (s.exc) return -1;
    if (!r) return 0;
    FakeObj *arr = O(r);
    if (arr->kind != OBJECTS || arr->elems.size() != 2 || !arr->elems[0] || !arr->elems[1]) { if (info) ++info[5]; return 1; }
    read_stats(O(arr->elems[0]), stats, info);
    read_stats(O(arr->elems[1]), stats ? stats + 5 : nullptr, info);
    for (int z = 0; z < d; ++z) {
        if (out_th) memcpy(out_th + (size_t)z * wh, O(o_th->elems[z])->floats.data(), wh * sizeof(float));
        if (out_sp) memcpy(out_sp + (size_t)z * wh, O(o_sp->elems[z])->floats.data(), wh * sizeof(float));
    }
    return 1;
}

}  // extern "C"

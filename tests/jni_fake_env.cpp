// jni_fake_env.cpp -- TEST ONLY.  A fake JNIEnv (arrays, classes, fields and a pending exception held in plain
// C++ objects) that drives the Java_... entry points of bonej2_b200/csrc/jni_shim.cpp, compiled against the test
// double tests/jni_stub/jni.h because the image has no JDK.  What it checks is the shim's own logic: every slice
// pinned once and released once (inputs with JNI_ABORT, outputs committed), nothing pinned when the call returns,
// null for a non-binary image, RuntimeException carrying bjt_last_error() on failure, MapStats fields set by name.
// tests/test_jni_shim.py calls the jt_* functions through ctypes.
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
void Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(JNIEnv *, jclass, jlong);
jobject Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(JNIEnv *, jclass, jlong, jobjectArray, jint, jint, jboolean,
                                                                     jboolean, jint, jdouble, jobjectArray);
}

struct _jfieldID { std::string name, sig; };

namespace {

enum Kind { CLASS, BYTES, FLOATS, OBJECTS, INSTANCE };

struct FakeObj : _jobject {
    Kind kind;
    std::string cls;                               // CLASS: its name; INSTANCE: the name of its class
    std::vector<uint8_t> bytes;
    std::vector<float> floats;
    std::vector<jobject> elems;
    std::map<std::string, long long> lfields;
    std::map<std::string, double> dfields;
    int pinned = 0;
    explicit FakeObj(Kind k) : kind(k) {}
};

struct State {
    std::vector<std::unique_ptr<FakeObj> > objs;
    std::vector<std::unique_ptr<_jfieldID> > fields;
    bool exc = false;
    std::string exc_class, exc_msg;
    int pins = 0, releases = 0, abort_releases = 0, commit_releases = 0, bad = 0;
    FakeObj *make(Kind k) { objs.emplace_back(new FakeObj(k)); return objs.back().get(); }
};

State &S(JNIEnv *e) { return *static_cast<State *>(e->fake_state); }
FakeObj *O(jobject o) { return static_cast<FakeObj *>(o); }

jclass f_FindClass(JNIEnv *e, const char *name) {
    const std::string n(name);
    if (n != "java/lang/RuntimeException" && n != "org/bonej/wrapperPlugins/b200/ThicknessNative$MapStats") return nullptr;
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
jsize f_GetArrayLength(JNIEnv *, jarray a) {
    FakeObj *o = O(a);
    return (jsize)(o->kind == BYTES ? o->bytes.size() : o->kind == FLOATS ? o->floats.size() : o->elems.size());
}
jobject f_GetObjectArrayElement(JNIEnv *e, jobjectArray a, jsize i) {
    FakeObj *o = O(a);
    if (o->kind != OBJECTS || i < 0 || (size_t)i >= o->elems.size()) { ++S(e).bad; return nullptr; }
    return o->elems[i];
}
void *f_GetPrimitiveArrayCritical(JNIEnv *e, jarray a, jboolean *is_copy) {
    FakeObj *o = O(a);
    if (is_copy) *is_copy = 0;
    ++S(e).pins; ++o->pinned;
    return o->kind == BYTES ? (void *)o->bytes.data() : o->kind == FLOATS ? (void *)o->floats.data() : nullptr;
}
void f_ReleasePrimitiveArrayCritical(JNIEnv *e, jarray a, void *p, jint mode) {
    FakeObj *o = O(a);
    State &s = S(e);
    const void *mine = o->kind == BYTES ? (const void *)o->bytes.data() : (const void *)o->floats.data();
    if (p != mine || o->pinned <= 0) ++s.bad;
    --o->pinned; ++s.releases;
    if (mode == JNI_ABORT) { ++s.abort_releases; if (o->kind != BYTES) ++s.bad; }      // only inputs may be discarded
    else if (mode == 0) { ++s.commit_releases; if (o->kind != FLOATS) ++s.bad; }
    else ++s.bad;
}

const BjtFakeJniTable TABLE = {f_FindClass, f_ThrowNew, f_AllocObject, f_GetFieldID, f_SetLongField, f_SetDoubleField,
                               f_GetArrayLength, f_GetObjectArrayElement, f_GetPrimitiveArrayCritical,
                               f_ReleasePrimitiveArrayCritical};

void copy_err(const State &s, char *err, int errlen) {
    if (!err || errlen <= 0) return;
    const std::string m = s.exc ? s.exc_class + ": " + s.exc_msg : std::string();
    strncpy(err, m.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
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

void jt_destroy(long long h) {
    State s;
    JNIEnv env{&TABLE, &s};
    Java_org_bonej_wrapperPlugins_b200_ThicknessNative_destroy(&env, nullptr, (jlong)h);
}

// ThicknessNative.thickness(...) on a byte[d][w*h] built from vol (every slice its own array, as ImageStack holds
// them).  Returns 1: a MapStats came back (stats = count, mean, stdDev, min, max; out filled when not null);
// 0: null without an exception; -1: an exception is pending (err).  info = pins, releases, input releases with
// JNI_ABORT, output releases committed, arrays still pinned afterwards, protocol violations.
int jt_thickness(long long h, const uint8_t *vol, int w, int hh, int d, int inverse, int mask, int threshold, double pixel_width,
                 float *out, double *stats, int *info, char *err, int errlen) {
    State s;
    JNIEnv env{&TABLE, &s};
    const size_t wh = (size_t)w * hh;
    FakeObj *in = s.make(OBJECTS), *outs = out ? s.make(OBJECTS) : nullptr;
    for (int z = 0; z < d; ++z) {
        FakeObj *a = s.make(BYTES);
        a->bytes.assign(vol + (size_t)z * wh, vol + (size_t)(z + 1) * wh);
        in->elems.push_back(a);
        if (outs) { FakeObj *f = s.make(FLOATS); f->floats.assign(wh, -1.0f); outs->elems.push_back(f); }
    }
    jobject r = Java_org_bonej_wrapperPlugins_b200_ThicknessNative_thickness(&env, nullptr, (jlong)h, in, w, hh, (jboolean)(inverse != 0),
                                                                             (jboolean)(mask != 0), threshold, pixel_width, outs);
    int still = 0;
    for (auto &o : s.objs) still += o->pinned != 0;
    if (info) { info[0] = s.pins; info[1] = s.releases; info[2] = s.abort_releases; info[3] = s.commit_releases; info[4] = still; info[5] = s.bad; }
    // the shim must not have written into the input arrays
    for (int z = 0; z < d; ++z)
        if (memcmp(O(in->elems[z])->bytes.data(), vol + (size_t)z * wh, wh) != 0 && info) ++info[5];
    copy_err(s, err, errlen);
    if (s.exc) return -1;
    if (!r) return 0;
    FakeObj *m = O(r);
    if (stats) {
        stats[0] = (double)m->lfields["count"]; stats[1] = m->dfields["mean"]; stats[2] = m->dfields["stdDev"];
        stats[3] = m->dfields["min"]; stats[4] = m->dfields["max"];
    }
    if (info && (m->lfields.size() != 1 || m->dfields.size() != 4)) ++info[5];
    if (out)
        for (int z = 0; z < d; ++z) memcpy(out + (size_t)z * wh, O(outs->elems[z])->floats.data(), wh * sizeof(float));
    return 1;
}

}  // extern "C"

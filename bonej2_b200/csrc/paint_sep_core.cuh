// paint_sep_core.cuh -- per-line logic of the separable sphere painter.  Plain C++ in
// __host__ __device__ functions so the CUDA kernels (paint_sep.cu) and the CPU model that tests
// the algorithm without a GPU (oracle/sep_paint_model.cpp) execute the same statements.
//
// Problem (Local_Thickness_Parallel): out(p) = max{ R_q : |p-q|^2 <= R_q } over the
// ridge points q.  With |p-q|^2 = dx^2 + dy^2 + dz^2 the ball test is peeled one axis per stage.
// An item (rho, R) at a voxel says "a ball tagged R reaches here with slack rho left for the
// remaining axes":
//   stage 1 (x)  ridge point (x_i, R)     -> item (R   - (x-x_i)^2, R) at x   while slack >= 0
//   stage 2 (y)  item (rho, R) at y_i     -> item (rho - (y-y_i)^2, R) at y   while slack >= 0
//   stage 3 (z)  out(z) = max R over items (rho, R) at z_i with (z-z_i)^2 <= rho
// Only the Pareto front matters at a voxel: (rho1,R1) makes (rho2,R2) redundant when rho1 >= rho2
// and R1 >= R2.  Every stage is a forward and a backward sweep along the line, each keeping an
// active set sorted by R descending.
//   stages 1-2: slack changes with position, so the active set is dynamic (SepActive).
//   stage 3: only max R is wanted, so dominance is static in (R, last covered z): a staircase with
//     R descending and reach ascending; the head is the answer (SepStair).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BJT_HD __host__ __device__ __forceinline__
#else
#define BJT_HD inline
#endif

// statistics hook of the CPU model's instrumented builds; compiles to nothing elsewhere
#ifndef BJT_SEP_COUNT
#define BJT_SEP_COUNT(k)
#endif
#ifndef BJT_SEP_TRACE
#define BJT_SEP_TRACE(kind, x, tc, dir)
#endif

namespace bjt {

struct SepItem { uint32_t rho, R; };

// one active item: tag R, slack rho0 when it arrived, arrival position t
struct SepEntry { uint32_t R, rho0; int32_t t; };

BJT_HD uint32_t sep_isqrt(uint32_t v) {
    // exact floor(sqrt(v)), v < 2^31
#if defined(__CUDA_ARCH__)
    uint32_t r = (uint32_t)sqrtf((float)v);
#else
    uint32_t r = (uint32_t)__builtin_sqrtf((float)v);
#endif
    while (r * r > v) --r;
    while ((r + 1u) * (r + 1u) <= v) ++r;
    return r;
}

// an integer >= floor(sqrt(v)) (and <= floor + 2): one conversion, one square root, no loop.
// A correctly rounded float sqrt is never below the largest integer <= the true root.
BJT_HD int sep_sqrt_upper(uint32_t v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)v));     // error of a few ulp: + 2 keeps it an upper bound
    return (int)r + 2;
#else
    return (int)__builtin_sqrtf((float)v) + 1;
#endif
}

// ---- storage policies --------------------------------------------------------------------------------
// A store holds two buffers ("halves") of up to cap() entries; step() reads one and writes the other.
template <int CAP>
struct SepLocalStore {                 // fixed arrays in the thread's frame
    SepEntry a[2][CAP];
    BJT_HD int cap() const { return CAP; }
    BJT_HD SepEntry load(int half, int i) const { return a[half][i]; }
    BJT_HD void store(int half, int i, const SepEntry &e) { a[half][i] = e; }
};

struct SepFlatStore {                  // a contiguous scratch region of 2 * capacity entries
    SepEntry *base;
    int capacity;
    BJT_HD int cap() const { return capacity; }
    BJT_HD SepEntry load(int half, int i) const { return base[(long long)half * capacity + i]; }
    BJT_HD void store(int half, int i, const SepEntry &e) { base[(long long)half * capacity + i] = e; }
};

struct SepUnpackPlain { static BJT_HD SepItem get(const SepItem *p, int k) { return p[k]; } };

// ---------------------------------------------------------------------------------------------
// dynamic active set (stages 1 and 2)
//
// One call of step() handles one position of the line: it merges the active set (sorted by R
// descending) with the items arriving here (two lists, each sorted by R descending) in a single
// pass, keeps a running "best slack so far" holder b, and for every visited item x decides:
//   slack_x <  0    : x has expired
//   slack_x >  best : x is on the front at this position (emitted), x becomes b
//   slack_x <= best : b dominates x now (R_b >= R_x).  x is dropped for good when b is not older
//                     than x (the gap then only grows) or when b still dominates x at (or beyond)
//                     x's own last position (the gap shrinks linearly, so it holds all the way);
//                     else x stays.
// Dropping needs a proof of redundancy, keeping never hurts: the result is exact.
template <class Store>
struct SepActive {
    int n, cur;
    Store st;

    BJT_HD void clear() { n = 0; cur = 0; }

    BJT_HD static int slack_at(int tt, int ti, uint32_t r0) {
        const int dt = tt - ti;
        return (int)r0 - dt * dt;              // |values| < 2^27 for extents <= 8192
    }

    // Returns the front size (may exceed fmax; only fmax are stored).  ok is cleared on capacity overflow.
    // REACH = false: emitted items are (slack, R), the Pareto front in slack.
    // REACH = true (the stage feeding stage 3, which only asks how far an item reaches): emitted items are
    //   (floor(sqrt(slack)), R) and an item is emitted only if it reaches strictly further than every
    //   emitted item with a larger tag -- equal reach with a smaller tag can never win there.
    template <class RawT, class Unpack, bool REACH>
    BJT_HD int step(int tc, int dir, bool skip_own, const RawT *lf, int nlf, const RawT *lb, int nlb, SepItem *out,
                    int fmax, bool &ok) {
        const int src = cur, dst = cur ^ 1;
        const int cap = st.cap();
        int i = 0, jf = 0, jb = 0, w = 0, nf = 0;
        int best = -1, best_t = 0, last_e = -1;
        uint32_t best_R = 0, best_rho0 = 0;
        bool last_emitted = false;
        SepItem af, ab;
        SepEntry ea;
        af.R = 0; af.rho = 0; ab.R = 0; ab.rho = 0; ea.R = 0; ea.rho0 = 0; ea.t = 0;
        if (nlf > 0) af = Unpack::get(lf, 0);
        if (nlb > 0) ab = Unpack::get(lb, 0);
        if (n > 0) ea = st.load(src, 0);
        while (true) {
            const uint32_t Ra = i < n ? ea.R : 0u;
            const uint32_t Rf = jf < nlf ? af.R : 0u, Rb = jb < nlb ? ab.R : 0u;
            if (!(Ra | Rf | Rb)) break;
            SepEntry x;
            bool fresh;
            if (Rf >= Ra && Rf >= Rb) { x.t = tc; x.rho0 = af.rho; x.R = Rf; fresh = true; ++jf; if (jf < nlf) af = Unpack::get(lf, jf); }
            else if (Rb >= Ra) { x.t = tc; x.rho0 = ab.rho; x.R = Rb; fresh = true; ++jb; if (jb < nlb) ab = Unpack::get(lb, jb); }
            else { x = ea; fresh = false; ++i; if (i < n) ea = st.load(src, i); }
            const int s = fresh ? (int)x.rho0 : slack_at(tc, x.t, x.rho0);
            BJT_SEP_COUNT(0);
            BJT_SEP_TRACE(i + jf + jb == 1 ? 0 : 1, x, tc, dir);
            if (s < 0) { BJT_SEP_COUNT(1); continue; }                     // expired
            if (s > best) {
                BJT_SEP_COUNT(2);
                const bool same_tag = best >= 0 && best_R == x.R && last_emitted;
                best = s; best_t = x.t; best_R = x.R; best_rho0 = x.rho0;
                const bool may_emit = !(skip_own && x.t == tc);
                if (!REACH) {
                    if (same_tag) --nf;                                    // same tag, more slack: replace
                    last_emitted = may_emit;
                    if (may_emit) {
                        if (nf < fmax) { out[nf].rho = (uint32_t)s; out[nf].R = x.R; }
                        ++nf;
                    }
                } else {
                    const int e = (int)sep_isqrt((uint32_t)s);
                    if (same_tag) {                                        // same tag: lengthen the entry in place
                        if (nf - 1 < fmax) out[nf - 1].rho = (uint32_t)e;
                        last_e = e;
                    } else if (e > last_e) {
                        last_e = e;                                        // an own arrival counts: the other sweep emits it
                        last_emitted = may_emit;
                        if (may_emit) {
                            if (nf < fmax) { out[nf].rho = (uint32_t)e; out[nf].R = x.R; }
                            ++nf;
                        }
                    } else {
                        last_emitted = false;
                    }
                }
            } else {
                if (dir > 0 ? best_t >= x.t : best_t <= x.t) { BJT_SEP_COUNT(3); continue; }   // dominated by an item that is not older
                const int ex = x.t + dir * sep_sqrt_upper(x.rho0);         // at or beyond x's last position
                if (slack_at(ex, best_t, best_rho0) >= slack_at(ex, x.t, x.rho0)) { BJT_SEP_COUNT(4); continue; }   // dominated for life
                BJT_SEP_COUNT(5);
                BJT_SEP_TRACE(2, x, tc, dir);
            }
            if (w >= cap) { ok = false; continue; }
            st.store(dst, w, x);
            ++w;
        }
        n = w; cur = dst;
        return nf;
    }
};

// ---------------------------------------------------------------------------------------------
// static staircase (stage 3): R descending, reach E (last covered coordinate in sweep direction) ascending.
// Store concept: cap(), getR(i), getE(i), put(i, R, E).
template <int CAP>
struct SepStairLocal {
    uint32_t R[CAP];
    int32_t E[CAP];
    BJT_HD int cap() const { return CAP; }
    BJT_HD uint32_t getR(int i) const { return R[i]; }
    BJT_HD int getE(int i) const { return E[i]; }
    BJT_HD void put(int i, uint32_t r, int e) { R[i] = r; E[i] = e; }
};

struct SepStairFlat {                  // scratch region: capacity R values, then capacity E values
    uint32_t *R;
    int32_t *E;
    int capacity;
    BJT_HD int cap() const { return capacity; }
    BJT_HD uint32_t getR(int i) const { return R[i]; }
    BJT_HD int getE(int i) const { return E[i]; }
    BJT_HD void put(int i, uint32_t r, int e) { R[i] = r; E[i] = e; }
};

template <class Store>
struct SepStair {
    int n;
    Store st;

    BJT_HD void clear() { n = 0; }

    // u = coordinate in sweep direction (z forward, -z backward); e = how far the item reaches.
    // Returns false on overflow.
    BJT_HD bool arrive(int u, uint32_t e, uint32_t Rn) {
        const int En = u + (int)e;
        // cheap reject against the head: R_n <= R_head and reach_n <= reach_head
        if (n > 0 && Rn <= st.getR(0) && En <= st.getE(0)) return true;
        int pos = 0;
        while (pos < n && st.getR(pos) > Rn) ++pos;
        if (pos > 0 && st.getE(pos - 1) >= En) return true;            // a larger tag reaches at least as far
        if (pos < n && st.getR(pos) == Rn && st.getE(pos) >= En) return true;
        int k = pos;
        while (k < n && st.getE(k) <= En) ++k;                          // R <= R_n and reach <= reach_n: redundant
        const int removed = k - pos;
        if (removed == 0) {
            if (n >= st.cap()) return false;
            for (int i = n; i > pos; --i) st.put(i, st.getR(i - 1), st.getE(i - 1));
            ++n;
        } else if (removed > 1) {
            const int sh = removed - 1;
            for (int i = k; i < n; ++i) st.put(i - sh, st.getR(i), st.getE(i));
            n -= sh;
        }
        st.put(pos, Rn, En);
        return true;
    }

    // value at coordinate u (after this position's arrivals): largest tag still reaching u
    BJT_HD uint32_t query(int u) {
        int h = 0;
        while (h < n && st.getE(h) < u) ++h;
        if (h > 0) {
            for (int i = h; i < n; ++i) st.put(i - h, st.getR(i), st.getE(i));
            n -= h;
        }
        return n > 0 ? st.getR(0) : 0u;
    }
};

}  // namespace bjt

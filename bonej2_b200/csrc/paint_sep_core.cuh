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

BJT_HD float sep_sqrt_fast(float v) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));            // a few ulp off at most
    return r;
#else
    return __builtin_sqrtf(v);
#endif
}

// exact floor(sqrt(v)), v < 2^31: the float root is off by far less than 1 (rounding of v: 2^-24 relative,
// the approximation: a few ulp), so its truncation is the answer or one beside it
BJT_HD uint32_t sep_isqrt(uint32_t v) {
    uint32_t r = (uint32_t)sep_sqrt_fast((float)v);
    if (r * r > v) --r;
    else if ((r + 1u) * (r + 1u) <= v) ++r;
    return r;
}

// an integer >= floor(sqrt(v)) (and <= floor + 2): one conversion, one square root, no loop.
// A correctly rounded float sqrt is never below the largest integer <= the true root.
BJT_HD int sep_sqrt_upper(uint32_t v) { return (int)sep_sqrt_fast((float)v) + 2; }

// ---- lanes of a warp as a patch of lines ----------------------------------------------------------------
// The lines of a stage form a plane: line = b * na + a (stage 1: a = y, b = z; stage 2: a = x, b = z;
// stage 3: a = x, b = y).  The lanes of a warp move in lock step, so a position costs the warp what its busiest
// line costs.  Lines that are close in BOTH directions of the plane have similar active sets; lines 31 apart
// in one direction sit in different structures.  Measured on the CPU model (tools/study/painter_study.cpp,
// 256^3 phantom, Tb.Sp): a warp of 8 x 4 lines needs 15.6 loop iterations per position of stage 2 where a
// warp of 32 x 1 needs 22.3 (stage 1: 3.9 against 5.6; stage 3: 6.7 against 8.5).  Eight lanes along `a` keep
// the 8-byte index words of a warp in 64-byte pieces.
// Thread `tid` of the grid -> its line, or -1 where the patch overhangs the plane (and for the threads a grid
// rounded up to whole blocks has beyond sep_patch_threads()).
constexpr int SEP_PATCH_A = 8, SEP_PATCH_B = 4;          // SEP_PATCH_A * SEP_PATCH_B == 32
BJT_HD long long sep_patch_warps(int na, int nb) {
    return (long long)((na + SEP_PATCH_A - 1) / SEP_PATCH_A) * ((nb + SEP_PATCH_B - 1) / SEP_PATCH_B);
}
BJT_HD long long sep_patch_threads(int na, int nb) { return sep_patch_warps(na, nb) * 32; }
BJT_HD long long sep_patch_line(long long tid, int na, int nb) {
    const long long warp = tid >> 5;
    const int lane = (int)(tid & 31);
    const long long pa = (na + SEP_PATCH_A - 1) / SEP_PATCH_A;
    const long long a = (warp % pa) * SEP_PATCH_A + lane % SEP_PATCH_A;
    const long long b = (warp / pa) * SEP_PATCH_B + lane / SEP_PATCH_A;
    return (a < na && b < nb) ? b * na + a : -1;
}

// ---- storage policies --------------------------------------------------------------------------------
// A store holds two buffers ("halves") of up to cap() entries; step() reads one and writes the other.
// Entries are opaque to step(): the policy names the type and its accessors, so a packed store moves
// one word around and never re-encodes it.
struct SepEntryOps {                   // plain struct entries
    typedef SepEntry Entry;
    static BJT_HD Entry make(uint32_t R, uint32_t rho0, int t) { Entry e; e.R = R; e.rho0 = rho0; e.t = t; return e; }
    static BJT_HD uint32_t R(const Entry &e) { return e.R; }
    static BJT_HD uint32_t rho0(const Entry &e) { return e.rho0; }
    static BJT_HD int t(const Entry &e) { return e.t; }
};

template <int CAP>
struct SepLocalStore : SepEntryOps {   // fixed arrays in the thread's frame
    SepEntry a[2][CAP];
    BJT_HD int cap() const { return CAP; }
    BJT_HD SepEntry load(int half, int i) const { return a[half][i]; }
    BJT_HD void store(int half, int i, const SepEntry &e) { a[half][i] = e; }
};

struct SepFlatStore : SepEntryOps {    // a contiguous scratch region of 2 * capacity entries
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
    typedef typename Store::Entry E;
    int n, cur;
    Store st;

    BJT_HD void clear() { n = 0; cur = 0; }

    // Returns the front size (may exceed fmax; only fmax are stored).  ok is cleared on capacity overflow.
    // REACH = false: emitted items are (slack, R), the Pareto front in slack.
    // REACH = true (the stage feeding stage 3, which only asks how far an item reaches): emitted items are
    //   (floor(sqrt(slack)), R) and an item is emitted only if it reaches strictly further than every
    //   emitted item with a larger tag -- equal reach with a smaller tag can never win there.
    // Visiting order: R descending; on equal R the forward list, then the backward list, then the active set.
    // DIR: sweep direction (+1 / -1); the backward sweep does not emit its own arrivals (the forward one did).
    template <class RawT, class Unpack, bool REACH, int DIR>
    BJT_HD int step(int tc, const RawT *lf, int nlf, const RawT *lb, int nlb, SepItem *out, int fmax, bool &ok) {
        constexpr int dir = DIR;
        constexpr bool skip_own = DIR < 0;
        const int src = cur, dst = cur ^ 1;
        const int cap = st.cap();
        int i = 0, jf = 0, jb = 0, w = 0, nf = 0;
        int best = -1, last_e = -1;
        bool last_emitted = false;
        const E none = Store::make(0u, 0u, tc);                            // R == 0: the source is exhausted
        E bx = none;                                                       // holder of `best`
        E ea = none, ef = none, eb = none;                                 // heads of the three sources
        if (n > 0) ea = st.load(src, 0);
        if (nlf > 0) { const SepItem a = Unpack::get(lf, 0); ef = Store::make(a.R, a.rho, tc); }
        if (nlb > 0) { const SepItem a = Unpack::get(lb, 0); eb = Store::make(a.R, a.rho, tc); }
        while (true) {
            const bool f_first = Store::R(ef) >= Store::R(eb);
            const E ev = f_first ? ef : eb;                                // head of the arrivals
            const uint32_t Ra = Store::R(ea), Rv = Store::R(ev);
            if (!(Ra | Rv)) break;
            E x;
            if (Rv >= Ra) {
                x = ev;
                if (f_first) {
                    ef = none;
                    if (++jf < nlf) { const SepItem a = Unpack::get(lf, jf); ef = Store::make(a.R, a.rho, tc); }
                } else {
                    eb = none;
                    if (++jb < nlb) { const SepItem a = Unpack::get(lb, jb); eb = Store::make(a.R, a.rho, tc); }
                }
            } else {
                x = ea;
                ea = none;
                if (++i < n) ea = st.load(src, i);
            }
            const int xt = Store::t(x);
            const uint32_t xr0 = Store::rho0(x);
            const int dt = tc - xt;
            const int s = (int)xr0 - dt * dt;                              // |values| < 2^27 for extents <= 8192
            BJT_SEP_COUNT(0);
            BJT_SEP_TRACE(i + jf + jb == 1 ? 0 : 1, x, tc, dir);
            if (s < 0) { BJT_SEP_COUNT(1); continue; }                     // expired
            if (s > best) {
                BJT_SEP_COUNT(2);
                const bool same_tag = best >= 0 && Store::R(bx) == Store::R(x) && last_emitted;
                best = s; bx = x;
                const bool may_emit = !(skip_own && xt == tc);
                if (!REACH) {
                    if (same_tag) --nf;                                    // same tag, more slack: replace
                    last_emitted = may_emit;
                    if (may_emit) {
                        if (nf < fmax) { out[nf].rho = (uint32_t)s; out[nf].R = Store::R(x); }
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
                            if (nf < fmax) { out[nf].rho = (uint32_t)e; out[nf].R = Store::R(x); }
                            ++nf;
                        }
                    } else {
                        last_emitted = false;
                    }
                }
            } else {
                const int bt = Store::t(bx);
                if (dir > 0 ? bt >= xt : bt <= xt) { BJT_SEP_COUNT(3); continue; }   // dominated by an item that is not older
                // dominated for life?  slack_b(u) - slack_x(u) is linear in u; it is >= 0 here, so test it at (or
                // beyond) x's last position ex:  (rho_b - rho_x) - (ex - bt)^2 + (ex - xt)^2 >= 0
                const int ex = xt + dir * sep_sqrt_upper(xr0);
                if ((int)Store::rho0(bx) - (int)xr0 + (bt - xt) * (2 * ex - bt - xt) >= 0) { BJT_SEP_COUNT(4); continue; }
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
    // hint: search start.  Every entry in front of it must have a tag > Rn; arrive() leaves the position it
    // found, so the items of one list (tags descending) can pass the same variable along, starting from 0.
    // Returns false on overflow.
    BJT_HD bool arrive(int u, uint32_t e, uint32_t Rn, int &hint) {
        const int En = u + (int)e;
        BJT_SEP_COUNT(6);
        // cheap reject against the head: R_n <= R_head and reach_n <= reach_head
        if (n > 0 && Rn <= st.getR(0) && En <= st.getE(0)) { BJT_SEP_COUNT(7); return true; }
        int pos = hint;
        while (pos < n && st.getR(pos) > Rn) { ++pos; BJT_SEP_COUNT(8); }
        hint = pos;
        if (pos > 0 && st.getE(pos - 1) >= En) { BJT_SEP_COUNT(9); return true; }   // a larger tag reaches at least as far
        if (pos < n && st.getR(pos) == Rn && st.getE(pos) >= En) { BJT_SEP_COUNT(9); return true; }
        int k = pos;
        while (k < n && st.getE(k) <= En) { ++k; BJT_SEP_COUNT(10); }   // R <= R_n and reach <= reach_n: redundant
        const int removed = k - pos;
        if (removed == 0) {
            if (n >= st.cap()) return false;
            for (int i = n; i > pos; --i) { st.put(i, st.getR(i - 1), st.getE(i - 1)); BJT_SEP_COUNT(11); }
            ++n;
        } else if (removed > 1) {
            const int sh = removed - 1;
            for (int i = k; i < n; ++i) { st.put(i - sh, st.getR(i), st.getE(i)); BJT_SEP_COUNT(11); }
            n -= sh;
        }
        st.put(pos, Rn, En);
        return true;
    }
    BJT_HD bool arrive(int u, uint32_t e, uint32_t Rn) { int hint = 0; return arrive(u, e, Rn, hint); }

    // value at coordinate u (after this position's arrivals): largest tag still reaching u
    BJT_HD uint32_t query(int u) {
        int h = 0;
        while (h < n && st.getE(h) < u) ++h;
        if (h > 0) {
            for (int i = h; i < n; ++i) { st.put(i - h, st.getR(i), st.getE(i)); BJT_SEP_COUNT(12); }
            n -= h;
        }
        return n > 0 ? st.getR(0) : 0u;
    }
};

}  // namespace bjt

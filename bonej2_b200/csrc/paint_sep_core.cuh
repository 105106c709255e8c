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

namespace bjt {

struct SepItem { uint32_t rho, R; };

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

// ---- storage policies: fixed arrays in the thread's frame, or strided views of a scratch buffer -----
template <int CAP, int HALVES = 2, int FIELDS = 4>
struct SepLocalStore {
    uint32_t a[HALVES][FIELDS][CAP];
    BJT_HD int cap() const { return CAP; }
    BJT_HD uint32_t &at(int half, int field, int i) { return a[half][field][i]; }
};

struct SepStridedStore {            // element (half, field, i) of a slot at base[((half*4+field)*cap + i) * stride]
    uint32_t *base;
    int capacity;
    long long stride;
    BJT_HD int cap() const { return capacity; }
    BJT_HD uint32_t &at(int half, int field, int i) { return base[((long long)(half * 4 + field) * capacity + i) * stride]; }
    static BJT_HD long long words_per_slot(int capacity_) { return 8ll * capacity_; }
};

struct SepUnpackPlain { static BJT_HD SepItem get(const SepItem *p, int k) { return p[k]; } };

// ---------------------------------------------------------------------------------------------
// dynamic active set (stages 1 and 2)
//
// One call of step() handles one position of the line: it merges the active set (sorted by R
// descending) with the items arriving here (two lists, each sorted by R descending) in a single
// pass, keeps a running "best slack so far" holder b, and for every visited item x decides:
//   slack_x >  best : x is on the front at this position (emitted), x becomes b
//   slack_x <= best : b dominates x now (R_b >= R_x).  x is dropped for good when b is not older
//                     than x (the gap then only grows) or when b still dominates x at x's own last
//                     position (the gap shrinks linearly, so it holds all the way); else x stays.
// Dropping needs a proof of redundancy, keeping never hurts: the result is exact.
template <class Store>
struct SepActive {
    enum { F_T = 0, F_E = 1, F_RHO = 2, F_R = 3 };
    int n, cur;
    Store st;

    BJT_HD void clear() { n = 0; cur = 0; }

    BJT_HD static int slack_at(int tt, int ti, uint32_t r0) {
        const int dt = tt - ti;
        return (int)r0 - dt * dt;              // |values| < 2^27 for extents <= 8192
    }

    // Returns the front size (may exceed fmax; only fmax are stored).  ok is cleared on capacity overflow.
    template <class RawT, class Unpack>
    BJT_HD int step(int tc, int dir, bool skip_own, const RawT *lf, int nlf, const RawT *lb, int nlb, SepItem *out,
                    int fmax, bool &ok) {
        const int src = cur, dst = cur ^ 1;
        const int cap = st.cap();
        int i = 0, jf = 0, jb = 0, w = 0, nf = 0;
        int best = -1, best_t = 0;
        uint32_t best_R = 0, best_rho0 = 0;
        bool last_emitted = false;
        SepItem af, ab;
        af.R = 0; af.rho = 0; ab.R = 0; ab.rho = 0;
        if (nlf > 0) af = Unpack::get(lf, 0);
        if (nlb > 0) ab = Unpack::get(lb, 0);
        while (true) {
            while (i < n && (dir > 0 ? (int)st.at(src, F_E, i) < tc : (int)st.at(src, F_E, i) > tc)) ++i;   // expired
            const uint32_t Ra = i < n ? st.at(src, F_R, i) : 0u;
            const uint32_t Rf = jf < nlf ? af.R : 0u, Rb = jb < nlb ? ab.R : 0u;
            if (!(Ra | Rf | Rb)) break;
            int ti, ei = 0;
            uint32_t r0, Ri;
            bool fresh;
            if (Rf >= Ra && Rf >= Rb) { ti = tc; r0 = af.rho; Ri = Rf; fresh = true; ++jf; if (jf < nlf) af = Unpack::get(lf, jf); }
            else if (Rb >= Ra) { ti = tc; r0 = ab.rho; Ri = Rb; fresh = true; ++jb; if (jb < nlb) ab = Unpack::get(lb, jb); }
            else { ti = (int)st.at(src, F_T, i); ei = (int)st.at(src, F_E, i); r0 = st.at(src, F_RHO, i); Ri = Ra; fresh = false; ++i; }
            const int s = fresh ? (int)r0 : slack_at(tc, ti, r0);
            if (s > best) {
                if (best >= 0 && best_R == Ri && last_emitted) --nf;      // same tag, more slack: replace
                best = s; best_t = ti; best_R = Ri; best_rho0 = r0;
                last_emitted = !(skip_own && ti == tc);
                if (last_emitted) {
                    if (nf < fmax) { out[nf].rho = (uint32_t)s; out[nf].R = Ri; }
                    ++nf;
                }
                if (fresh) ei = tc + dir * (int)sep_isqrt(r0);
            } else {
                if (dir > 0 ? best_t >= ti : best_t <= ti) continue;       // dominated by an item that is not older
                if (fresh) ei = tc + dir * (int)sep_isqrt(r0);
                if (slack_at(ei, best_t, best_rho0) >= slack_at(ei, ti, r0)) continue;   // dominated for life
            }
            if (w >= cap) { ok = false; continue; }
            st.at(dst, F_T, w) = (uint32_t)ti; st.at(dst, F_E, w) = (uint32_t)ei; st.at(dst, F_RHO, w) = r0; st.at(dst, F_R, w) = Ri;
            ++w;
        }
        n = w; cur = dst;
        return nf;
    }
};

// ---------------------------------------------------------------------------------------------
// static staircase (stage 3): R descending, reach E (last covered coordinate in sweep direction) ascending
template <class Store>      // uses fields 0 = R, 1 = E of half 0
struct SepStair {
    int n;
    Store st;

    BJT_HD void clear() { n = 0; }
    BJT_HD uint32_t &Rr(int i) { return st.at(0, 0, i); }
    BJT_HD uint32_t &Er(int i) { return st.at(0, 1, i); }

    // u = coordinate in sweep direction (z forward, -z backward).  Returns false on overflow.
    BJT_HD bool arrive(int u, uint32_t rho, uint32_t Rn) {
        // cheap reject against the head: R_n <= R_head and reach_n <= reach_head
        if (n > 0 && Rn <= Rr(0)) {
            const int span = (int)Er(0) - u;
            if (span >= 0 && rho < (uint32_t)(span + 1) * (uint32_t)(span + 1)) return true;
        }
        const int En = u + (int)sep_isqrt(rho);
        int pos = 0;
        while (pos < n && Rr(pos) > Rn) ++pos;
        if (pos > 0 && (int)Er(pos - 1) >= En) return true;              // a larger tag reaches at least as far
        if (pos < n && Rr(pos) == Rn && (int)Er(pos) >= En) return true;
        int k = pos;
        while (k < n && (int)Er(k) <= En) ++k;                            // R <= R_n and reach <= reach_n: redundant
        const int removed = k - pos;
        if (removed == 0) {
            if (n >= st.cap()) return false;
            for (int i = n; i > pos; --i) { Rr(i) = Rr(i - 1); Er(i) = Er(i - 1); }
            ++n;
        } else if (removed > 1) {
            const int sh = removed - 1;
            for (int i = k; i < n; ++i) { Rr(i - sh) = Rr(i); Er(i - sh) = Er(i); }
            n -= sh;
        }
        Rr(pos) = Rn; Er(pos) = (uint32_t)En;
        return true;
    }

    // value at coordinate u (after this position's arrivals): largest tag still reaching u
    BJT_HD uint32_t query(int u) {
        int h = 0;
        while (h < n && (int)Er(h) < u) ++h;
        if (h > 0) {
            for (int i = h; i < n; ++i) { Rr(i - h) = Rr(i); Er(i - h) = Er(i); }
            n -= h;
        }
        return n > 0 ? Rr(0) : 0u;
    }
};

}  // namespace bjt

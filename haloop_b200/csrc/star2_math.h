// star2_math.h — the lane arithmetic of the fused star-CTC path (star2.cuh), written as host + device functions so
// that tools/star2_host_check.cpp can run the very same code on the CPU, lane by lane, against the float64 oracle.
//
// State layout (ha/star.py:91-145): position k of the target owns the quad
//     b0 = blank (j = 4k)   st = star "anything but y_k" (4k+1)   b1 = blank (4k+2)   lb = label y_k (4k+3)
// and the final quad k = L holds (blank, last star, final blank) only.  Transitions forward in time:
//     b0 <- lb[k-1], b0        st <- b0, st, b1 (x star_penalty)        b1 <- st, b1
//     lb <- b0, st, b1, and lb[k-1] unless y_k == y_{k-1}               (labels have no self loop)
//
// Numbers: "quad-normalised" linear domain, the ctc2.cuh format with four states per exponent.  A lane owns J (4, 2 or 1)
// consecutive quads in POSITION order (component c is position J g + c - 4); the four states of a quad are plain
// fp32 values sharing one int32 exponent, rescaled after every step so that the largest sits in [2^32, 2^33).
#pragma once
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define S2_HD __host__ __device__ __forceinline__
#else
#define S2_HD inline
#endif

namespace hab {

constexpr int kQLaneExp = 32 + 127;       // biased exponent the largest state of a quad is normalised to
constexpr int kQAlignMax = 30;            // largest up-shift applied to a neighbour's state
constexpr int kQVoidE = -(1 << 28);       // exponent of an all-zero quad (common.cuh kVoidE)
constexpr float kQMinNormal = 1.1754943508222875e-38f;
constexpr float kQLog2e = 1.4426950408889634f;
constexpr float kQMagic = 12582912.0f;

S2_HD int s2_f2i(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(x);
#else
    int i; memcpy(&i, &x, 4); return i;
#endif
}
S2_HD float s2_i2f(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float x; memcpy(&x, &i, 4); return x;
#endif
}
S2_HD float s2_ex2(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return exp2f(x);
#endif
}
S2_HD float s2_rcp(float x) {             // 1 / x for a positive normal x (1 ulp: MUFU.RCP)
#if defined(__CUDA_ARCH__)
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / x;
#endif
}
S2_HD int s2_min(int a, int b) { return a < b ? a : b; }
S2_HD int s2_max(int a, int b) { return a > b ? a : b; }

// p = 2^(x log2e - l2), the emission2() of ctc2.cuh: l2 split into a multiple of 2^-10 and a remainder so that the
// integer part of the exponent is removed before the one rounding that matters.  Floor 2^-125.75.
struct QRowNorm { float l2q, dl; };
S2_HD QRowNorm s2_row_norm(float l2) {
    QRowNorm r;
    r.l2q = rintf(l2 * 1024.0f) * (1.0f / 1024.0f);
    r.dl = l2 - r.l2q;
    return r;
}
S2_HD float s2_emission(float x, QRowNorm rn) {
    const float t = fmaxf(fmaf(x, kQLog2e, -rn.l2q), -125.0f);
    const float tk = t + kQMagic;
    const float kf = tk - kQMagic;
    const float f = fmaxf(fmaf(x, kQLog2e, -(rn.l2q + kf)) - rn.dl, -0.75f);
    return s2_i2f(s2_f2i(s2_ex2(f)) + (int)((unsigned)s2_f2i(tk) << 23));
}
// Star emission P - p_y (ha/star.py:4-5 logsubexp, :32) in the linear domain: (s_nb - e_y) * pscale, with s_nb the
// sum of the non-blank terms 2^(x log2e - m2) and e_y the very term class y contributed to it, so the difference is
// the sum over the other classes up to the rounding of s_nb.  y == 0: the all-star P (ha/star.py:31).
S2_HD float s2_star_emission(float x, bool excludes, float s_nb, float m2, float pscale) {
    const float ey = excludes ? s2_ex2(fmaf(x, kQLog2e, -m2)) : 0.0f;
    return fmaxf((s_nb - ey) * pscale, kQMinNormal);
}

// x * 2^d for x >= 0, -512 <= d <= 64: one exact multiplication by 2^d (zeros stay zero — exponent-field arithmetic on a
// zero would create 2^(d-127) out of nothing); a neighbour more than 2^126 below is dropped (it is < 2^-124 of the quad's
// largest state).
S2_HD float s2_scale_pow2(float x, int d) {
    return (d >= -126) ? x * s2_i2f((d + 127) << 23) : 0.0f;
}

template <int J> struct QLane { float b0[J], st[J], b1[J], lb[J]; int e[J]; };
template <int J> struct QSums { float w0[J], vs[J], u1[J], vl[J]; int e[J]; };      // pre-emission sums of b0, st, b1, lb; scale e[c] (= s.e[c] on return)

template <int J>
S2_HD void s2_lane_clear(QLane<J>& s) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < J; ++c) { s.b0[c] = 0.0f; s.st[c] = 0.0f; s.b1[c] = 0.0f; s.lb[c] = 0.0f; s.e[c] = kQVoidE; }
}

// First half of a step.  DIR 0 (alpha, forward in time): nl / ne = the label state of the quad below my lowest one
// and its exponent.  DIR 1 (beta): n0 / nl / ne = the first blank and the label of the quad above my highest one.
//   alpha   w0 = b0 + lb[k-1]       vs = b0 + st + b1            u1 = st + b1      vl = vs + [allowed] lb[k-1]
//   beta    w0 = b0 + st + lb       vs = u1 = st + b1 + lb                         vl = b0[k+1] + [allowed] lb[k+1]
// (ha/star.py:123-145; beta: the transposed arcs).  A quad dwarfed by its neighbour (the wavefront arrives: exponent
// distance d > kQAlignMax) moves its exponent up by sh = d - kQAlignMax: its own sums are multiplied by fo = 2^-sh.
// This is BRANCH-FREE — fo is exactly 1 in all but a few steps of an utterance — because a divergent branch around the
// rare case costs far more than the eight instructions per quad: the convergence barrier it needs splits the step
// into blocks the scheduler cannot interleave (measured: 0.161 -> 0.131 ms for the forward kernel at one CTA per SM).
template <int DIR, int J>
S2_HD void s2_quad_sums(QLane<J>& s, unsigned allowed, float n0, float nl, int ne, QSums<J>& q) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < J; ++c) {
        const bool edge = DIR ? (c == J - 1) : (c == 0);
        const int cb = DIR ? (c < J - 1 ? c + 1 : c) : (c ? c - 1 : 0);
        const float xl = edge ? nl : s.lb[cb];
        const float x0 = DIR ? (edge ? n0 : s.b0[cb]) : 0.0f;
        const int d = (edge ? ne : s.e[cb]) - s.e[c];            // neighbour's exponent above mine
        const int sh = s2_max(d - kQAlignMax, 0);                // > 0: I move up
        const int dd = s2_max(d - sh, -512);                     // the neighbour's shift: <= kQAlignMax
        const float fo = (sh <= 126) ? s2_i2f((127 - sh) << 23) : 0.0f;      // 2^-sh (1 when sh == 0)
        const bool al = (allowed >> c) & 1u;
        const float cl = s2_scale_pow2(xl, dd);
        q.e[c] = s.e[c] + sh;
        if (DIR == 0) {
            const float u = s.st[c] + s.b1[c];
            const float v = u + s.b0[c];
            q.w0[c] = fmaf(s.b0[c], fo, cl); q.vs[c] = v * fo; q.u1[c] = u * fo; q.vl[c] = fmaf(v, fo, al ? cl : 0.0f);
        } else {
            const float c0 = s2_scale_pow2(x0, dd);
            const float x = s.st[c] + s.lb[c];
            const float z = (s.b1[c] + x) * fo;
            q.w0[c] = (s.b0[c] + x) * fo; q.vs[c] = z; q.u1[c] = z; q.vl[c] = c0 + (al ? cl : 0.0f);
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < J; ++c) s.e[c] = q.e[c];                 // (after every component has read its neighbour's old exponent)
}

// Second half: multiply by the emissions (pb blank, pl[c] label, ps[c] star, pen = exp(star_penalty), paid on every
// arc into a star, ha/star.py:137) and renormalise each quad.
template <int J>
S2_HD void s2_quad_emit(QLane<J>& s, const QSums<J>& q, float pb, const float (&pl)[J], const float (&ps)[J], float pen) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < J; ++c) {
        const float n0 = q.w0[c] * pb, ns = (q.vs[c] * ps[c]) * pen, n1 = q.u1[c] * pb, nl = q.vl[c] * pl[c];
        const float mx = fmaxf(fmaxf(n0, ns), fmaxf(n1, nl));
        const int delta = s2_min(kQLaneExp - (s2_f2i(mx) >> 23), 120);
        const float f = s2_i2f((delta + 127) << 23);
        s.b0[c] = n0 * f; s.st[c] = ns * f; s.b1[c] = n1 * f; s.lb[c] = nl * f;
        s.e[c] = (mx > 0.0f) ? s.e[c] - delta : kQVoidE;
    }
}

// Occupancies of a frame from my pre-emission sums and the OTHER side's stored (emission included) label and star
// states of the same frame: gl = gamma(label k), gs = gamma(star k), h = gamma(star k) / (P - p_{y_k}).
// eZ / rZ: exponent of Z and 1 / mantissa of Z.
template <int J>
S2_HD void s2_quad_occ(const QLane<J>& s, const QSums<J>& q, const float (&lbo)[J], const float (&sto)[J], const int (&eo)[J],
                       int eZ, float rZ, const float (&ps)[J], float (&gl)[J], float (&gs)[J], float (&h)[J]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < J; ++c) {
        const int xs = s.e[c] + eo[c] - eZ;
        const float sc = (xs < -126) ? 0.0f : s2_i2f(s2_f2i(rZ) + (int)((unsigned)s2_min(xs, 90) << 23));
        gl[c] = (q.vl[c] * lbo[c]) * sc;
        gs[c] = (q.vs[c] * sto[c]) * sc;
        h[c] = (ps[c] > 0.0f) ? gs[c] * s2_rcp(ps[c]) : 0.0f;      // (an IEEE division takes its slow path on the tiny occupancies)
    }
}

// skip-transition bits of a lane's J quads (group g, labels y[0..L)), by component   [ha/star.py:117-118, :139-140]
template <int J>
S2_HD unsigned s2_allowed(int g, int dir, int L, const int* y, int label_mask) {
    unsigned allowed = 0;
    for (int c = 0; c < J; ++c) {
        const int p = J * g + c - 4;                 // target position of the quad
        bool al = false;
        if (!dir) {
            if (p == 0) al = true;                   // the virtual source below the first quad
            else if (p >= 1 && p < L) al = (y[p] & label_mask) != (y[p - 1] & label_mask);
        } else {
            if (p == L - 1) al = true;               // the virtual source above the final quad
            else if (p >= 0 && p + 1 < L) al = (y[p + 1] & label_mask) != (y[p] & label_mask);
        }
        if (al) allowed |= 1u << c;
    }
    return allowed;
}

}  // namespace hab

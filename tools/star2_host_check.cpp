// star2_host_check.cpp — CPU emulation of the fused star-CTC kernels (haloop_b200/csrc/star2.cuh): the lane
// arithmetic is the shipped code (star2_math.h compiled for the host), the lanes of a side are stepped one after the
// other from a snapshot of the previous frame, rows / boundaries / occupancy slots use the kernel's index formulas.
// Built and run by tools/star2_host_check.py against the float64 oracle.  Test infrastructure, not product code.
//   g++ -O1 -shared -fPIC -o /tmp/libstar2_host.so tools/star2_host_check.cpp
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "../haloop_b200/csrc/star2_math.h"

using namespace hab;

namespace {

template <int J>
struct Side {
    int dir, NL;
    std::vector<QLane<J>> s;
    std::vector<unsigned> allowed;
    std::vector<int> g;
};

inline int na_of(int L) { return (L + 5 + 3) / 4 * 4; }      // label indices a = position + 4 of an utterance, padded to 4

template <int J>
void side_init(Side<J>& sd, int dir, int L, const int* y) {
    sd.dir = dir;
    sd.NL = na_of(L) / J;
    sd.s.resize(sd.NL);
    sd.allowed.resize(sd.NL);
    sd.g.resize(sd.NL);
    for (int gl = 0; gl < sd.NL; ++gl) {
        const int g = dir ? sd.NL - 1 - gl : gl;
        sd.g[gl] = g;
        sd.allowed[gl] = s2_allowed<J>(g, dir, L, y, 0x3fffffff);
        s2_lane_clear(sd.s[gl]);
    }
    // the virtual source: mass 1 on the label below the first quad (alpha) / above the final quad (beta)
    const int a_inj = dir ? L + 4 : 3;
    const int gi = a_inj / J, ci = a_inj % J;
    const int gl = dir ? sd.NL - 1 - gi : gi;
    sd.s[gl].lb[ci] = 1.0f; sd.s[gl].e[ci] = 0;
}

// neighbour values entering lane gl of a side (the kernel: shuffle from lane - 1 / mailbox of the warp below)
template <int J>
void fetch(const Side<J>& sd, const std::vector<QLane<J>>& old, int gl, float& n0, float& nl, int& ne) {
    if (gl == 0) { n0 = 0.0f; nl = 0.0f; ne = kQVoidE; return; }
    const QLane<J>& p = old[gl - 1];
    const int c = sd.dir ? 0 : J - 1;
    n0 = p.b0[c]; nl = p.lb[c]; ne = p.e[c];
}

struct Frame { float pb; std::vector<float> pl, ps; float l2; };

}  // namespace

template <int J>
int star2_host_j(const float* x, int T, int V, const int* y, int S, int L, int Tn, float star_penalty,
                 int from_logits, float gout, float* loss_out, float* grad /* [T][V] */) {
    const int NA = na_of(L), NL = NA / J;
    const float pen = expf(star_penalty);
    // ---- rows: statistics + gather (the row warps) ----
    std::vector<Frame> fr(Tn);
    for (int t = 0; t < Tn; ++t) {
        const float* row = x + (size_t)t * V;
        float s_nb = 0.0f;
        for (int c = 1; c < V; ++c) s_nb += s2_ex2(row[c] * kQLog2e);
        const float e0 = s2_ex2(row[0] * kQLog2e);
        float m2 = 0.0f;
        const float tot = s_nb + e0;
        const float l2 = from_logits ? m2 + log2f(tot) : 0.0f;
        const float pscale = from_logits ? 1.0f / tot : s2_ex2(m2);
        const QRowNorm rn = s2_row_norm(l2);
        Frame& f = fr[t];
        f.l2 = l2;
        f.pb = s2_emission(row[0], rn);
        f.pl.assign(NA, 0.0f); f.ps.assign(NA, 0.0f);
        for (int k = 0; k <= L; ++k) {
            const int yk = (k < S) ? y[k] : 0;
            f.pl[k + 4] = (k < L) ? s2_emission(row[yk], rn) : 0.0f;
            f.ps[k + 4] = s2_star_emission(row[yk], yk != 0, s_nb, m2, pscale);
        }
    }
    Side<J> sd[2];
    side_init(sd[0], 0, L, y);
    side_init(sd[1], 1, L, y);
    const int tm = Tn >> 1;
    // stored rows: [lb 4NL][st 4NL][e 4NL] per frame, position order
    std::vector<float> tr_lb((size_t)Tn * NA), tr_st((size_t)Tn * NA);
    std::vector<int> tr_e((size_t)Tn * NA);

    auto step = [&](Side<J>& S_, int t, bool phase2, std::vector<float>* GL, std::vector<float>* GS, std::vector<float>* H,
                    int eZ, float rZ) {
        std::vector<QLane<J>> old = S_.s;
        for (int gl = 0; gl < S_.NL; ++gl) {
            float n0, nl; int ne;
            fetch(S_, old, gl, n0, nl, ne);
            QLane<J>& s = S_.s[gl];
            const int g = S_.g[gl];
            QSums<J> q;
            if (S_.dir) s2_quad_sums<1>(s, S_.allowed[gl], n0, nl, ne, q); else s2_quad_sums<0>(s, S_.allowed[gl], n0, nl, ne, q);
            float pl[J], ps[J];
            for (int c = 0; c < J; ++c) { pl[c] = fr[t].pl[J * g + c]; ps[c] = fr[t].ps[J * g + c]; }
            if (phase2) {
                float lbo[J], sto[J], gl4[J], gs4[J], h4[J]; int eo[J];
                for (int c = 0; c < J; ++c) {
                    lbo[c] = tr_lb[(size_t)t * NA + J * g + c]; sto[c] = tr_st[(size_t)t * NA + J * g + c];
                    eo[c] = tr_e[(size_t)t * NA + J * g + c];
                }
                s2_quad_occ(s, q, lbo, sto, eo, eZ, rZ, ps, gl4, gs4, h4);
                for (int c = 0; c < J; ++c) { (*GL)[J * g + c] = gl4[c]; (*GS)[J * g + c] = gs4[c]; (*H)[J * g + c] = h4[c]; }
            }
            s2_quad_emit(s, q, fr[t].pb, pl, ps, pen);
            if (!phase2) {
                for (int c = 0; c < J; ++c) {
                    tr_lb[(size_t)t * NA + J * g + c] = s.lb[c]; tr_st[(size_t)t * NA + J * g + c] = s.st[c];
                    tr_e[(size_t)t * NA + J * g + c] = s.e[c];
                }
            }
        }
    };
    for (int i = 0; i < tm; ++i) step(sd[0], i, false, nullptr, nullptr, nullptr, 0, 0.0f);
    for (int i = 0; i < Tn - tm; ++i) step(sd[1], Tn - 1 - i, false, nullptr, nullptr, nullptr, 0, 0.0f);

    // ---- the meeting: Z = sum over the states of (alpha's pre-emission sums of frame tm) x (beta's boundary) ----
    // boundaries in position order
    std::vector<QLane<J>> ba(NL), bb(NL);
    for (int gl = 0; gl < NL; ++gl) { ba[sd[0].g[gl]] = sd[0].s[gl]; bb[sd[1].g[gl]] = sd[1].s[gl]; }
    double tot = 0.0; int pm = 4 * kQVoidE;
    std::vector<float> zm; std::vector<int> zx;
    {
        std::vector<QLane<J>> sa = ba;
        for (int g = 0; g < NL; ++g) {
            float nl = 0.0f; int ne = kQVoidE;
            if (g > 0) { nl = ba[g - 1].lb[J - 1]; ne = ba[g - 1].e[J - 1]; }
            QSums<J> q;
            s2_quad_sums<0>(sa[g], s2_allowed<J>(g, 0, L, y, 0x3fffffff), 0.0f, nl, ne, q);
            for (int c = 0; c < J; ++c) {
                const float m[4] = {q.w0[c] * bb[g].b0[c], q.vs[c] * bb[g].st[c], q.u1[c] * bb[g].b1[c], q.vl[c] * bb[g].lb[c]};
                for (int j = 0; j < 4; ++j) {
                    zm.push_back(m[j]);
                    const int ex = (m[j] > 0.0f) ? sa[g].e[c] + bb[g].e[c] + (s2_f2i(m[j]) >> 23) : 4 * kQVoidE;
                    zx.push_back(ex);
                    pm = s2_max(pm, ex);
                }
            }
        }
        for (size_t j = 0; j < zm.size(); ++j)
            if (zm[j] > 0.0f) {
                const float mant = s2_i2f((s2_f2i(zm[j]) & 0x007fffff) | 0x3f800000);
                const int rel = zx[j] - pm;
                if (rel > -1000) tot += scalbn((double)mant, rel);
            }
    }
    if (!(pm > -(1 << 27)) || !(tot > 0.0)) { *loss_out = INFINITY; return 0; }
    const int ex = ilogb(tot);
    const double log2z = (double)(pm - 127) + log2(tot);
    *loss_out = (float)(-log2z * 0.6931471805599453094);
    const int eZ = pm - 127 + ex;
    const float rZ = (float)(1.0 / scalbn(tot, -ex));

    // ---- phase 2: occupancies + gradient rows ----
    const float delta = from_logits ? 1.0f : 0.0f;
    for (int t = 0; t < T; ++t)
        for (int c = 0; c < V; ++c) grad[(size_t)t * V + c] = 0.0f;
    std::vector<float> GL(NA), GS(NA), H(NA);
    auto grad_row = [&](int t) {
        const float* row = x + (size_t)t * V;
        float* g = grad + (size_t)t * V;
        for (int c = 0; c < V; ++c) g[c] = gout * s2_ex2(fmaf(row[c], kQLog2e, -fr[t].l2));        // pre(): g * p_c
        float bs = 0.0f, G = 0.0f;
        std::vector<float> a(L + 1);
        for (int k = 0; k <= L; ++k) {
            const int yk = (k < S) ? y[k] : 0;
            const float gl = GL[k + 4], gs = GS[k + 4], h = H[k + 4];
            bs += gl + gs; G += h;
            a[k] = ((yk != 0) ? g[yk] * h : 0.0f) - gout * gl;
        }
        const float r0 = g[0];
        for (int c = 0; c < V; ++c) g[c] *= (delta - G);
        g[0] = r0 * delta - gout * (1.0f - bs);
        for (int k = 0; k <= L; ++k) { const int yk = (k < S) ? y[k] : 0; g[yk] += a[k]; }
    };
    for (int i = tm; i < Tn; ++i) { step(sd[0], i, true, &GL, &GS, &H, eZ, rZ); grad_row(i); }
    for (int i = Tn - tm; i < Tn; ++i) { const int t = Tn - 1 - i; step(sd[1], t, true, &GL, &GS, &H, eZ, rZ); grad_row(t); }
    return 0;
}

extern "C" int star2_host(const float* x, int T, int V, const int* y, int S, int L, int Tn, float star_penalty,
                          int from_logits, float gout, float* loss_out, float* grad, int J) {
    if (J == 1) return star2_host_j<1>(x, T, V, y, S, L, Tn, star_penalty, from_logits, gout, loss_out, grad);
    if (J == 2) return star2_host_j<2>(x, T, V, y, S, L, Tn, star_penalty, from_logits, gout, loss_out, grad);
    return star2_host_j<4>(x, T, V, y, S, L, Tn, star_penalty, from_logits, gout, loss_out, grad);
}

// star2.cuh — fused star / wildcard CTC (STC) loss + logit gradient for sm_100a.  Replaces ha/star.py:65-163
// (star_ctc_forward_score), :8-49 (intersperse_stars: the (T,N,2V) tensor is never built) and the autograd backward
// with the two-kernel shape of ctc2.cuh — emissions and occupancies never reach HBM:
//
//   star2_fwd_kernel  one CTA per (utterance, sweep direction).  Row warps stream the logit rows of their side's half
//                     of the frames through a TMA ring, form the row statistics (log-sum-exp, the non-blank sum P) and
//                     gather blank, label and star emissions (P - p_y, ha/star.py:32) straight into a shared-memory
//                     slot; the trellis warps consume the slots, advance alpha (direction 0) or beta (direction 1) and
//                     store the LABEL and STAR states of each frame.  The sides meet in the middle; whichever CTA
//                     finishes second forms Z from both boundaries.
//   star2_bwd_kernel  the other half of each side's frames: row warps reload the logit row, gather the emissions
//                     again, the trellis warps advance the recursion and multiply their pre-emission sums with the rows
//                     the OTHER side stored, leaving label occupancies, star occupancies and h_k = gamma(star k) /
//                     (P - p_{y_k}) in the slot; the row warp that owns the frame forms the dense gradient
//                     g (p_c (delta - G + H_c) - occ_c) in place and bulk-stores it.
//
// Lane arithmetic: star2_math.h (quad-normalised linear domain; checked on the CPU by tools/star2_host_check.py).
#pragma once
#include <type_traits>

#include "common.cuh"
#include "ctc2.cuh"
#include "star2_math.h"

namespace hab {

// -DHAB_STAR2_PROBE: clock64 probes of one CTA pair (tools/star2_probe.py reads them from the workspace header)
#ifdef HAB_STAR2_PROBE
#define S2P_DECL(n) long long s2p_acc[n] = {}; long long s2p_t = clock64(); const bool s2p_on = (blockIdx.x < 2) && (lane == 0)
#define S2P_MARK(i) do { const long long s2p_n = clock64(); s2p_acc[i] += s2p_n - s2p_t; s2p_t = s2p_n; } while (0)
#define S2P_DUMP(base, n, cnt) do { if (s2p_on) { for (int s2p_i = 0; s2p_i < (n); ++s2p_i) p.hdr[(base) + s2p_i] = (float)s2p_acc[s2p_i] / (float)max((cnt), 1); } } while (0)
#else
#define S2P_DECL(n)
#define S2P_MARK(i)
#define S2P_DUMP(base, n, cnt)
#endif

// R row warps per CTA (template parameter, 4 or 8); 2 R emission / occupancy slots (a row warp gathers its next row
// while the occupancies of its previous one are still in their slot)

struct Star2Ws {                  // workspace layout (byte offsets); the 256-byte header holds exp(star_penalty)
    size_t meta, order, tgt, nflist, nfhdr, loss, zinfo, cnt, stat, bound, scr, tr, total;
    int Sp, NLmax, SPL, BW, NF;
};
__host__ inline Star2Ws star2_ws_layout(int T, int N, int S) {
    Star2Ws w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.NLmax = S / 4 + 2;                        // lanes (groups of 4 quads: positions -4 .. S) per side
    w.SPL = 12 * w.NLmax;                       // stored row: [4 NL labels][4 NL stars][4 NL quad exponents]
    w.BW = 20 * w.NLmax;                        // boundary: [b0][st][b1][lb][exponents], 4 NL each
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.NF = w.Sp + 32 * kRcap;
    w.nflist = take(sizeof(int) * (size_t)N * w.NF);
    w.nfhdr = take(sizeof(int2) * (size_t)N);
    w.loss = take(sizeof(float) * (size_t)N);
    w.zinfo = take(sizeof(int4) * (size_t)N);
    w.cnt = take(sizeof(int) * (size_t)N);
    w.stat = take(sizeof(float4) * (size_t)N * T);      // per frame {l2, non-blank sum, shift m2, star scale}
    w.bound = take(sizeof(int) * (size_t)N * 2 * w.BW);
    w.scr = take(sizeof(int) * (size_t)N * 2 * 32 * 16);   // 64 bytes per lane of a side's last warp: where lanes without a group store
    w.tr = take(sizeof(int) * (size_t)N * T * w.SPL);
    w.total = o;
    return w;
}

struct Star2Params {
    const float* x; long long sx_t, sx_n;
    float* gx; long long sg_t, sg_n;
    int T, N, V, S, Sp;
    const int4* meta; const int* order; const int* tgt; const int* nflist; const int2* nfhdr; int NF;
    float4* stat; int* tr; int SPL; int* bound; int BW; int* scr; int4* zinfo; int* cnt;
    float* loss; float* loss_ws; const float* gout; float* hdr;
    float pen;                            // exp(star_penalty) (forward; the backward reads it from the header)
    int from_logits, NS, NLmax, NA, EMF;  // ring stages per row warp; lanes of the longest target; 4 NLmax; floats per slot
};

// A slot: [0] blank; label index a = position + 4 (a < 4 and a > L + 4 are phantoms that stay 0):
//   [4 + a] label emission -> label occupancy     [4 + NA + a] star emission -> h     [4 + 2 NA + a] -> star occupancy
__host__ __device__ inline int star2_em_floats(int NLmax) { return 4 + 12 * NLmax; }

struct Star2Smem { int bars, mail, red, tscr, tgt, em, st, rows, total; };
__host__ __device__ inline Star2Smem star2_smem(int W, int R, int NS, int V, int Sp, int NLmax, bool bwd) {
    const int kSR = R, kSNE = 2 * R;
    Star2Smem s;
    int o = 0;
    auto take = [&](int bytes) { int at = o; o = round_up(o + bytes, 128); return at; };
    s.bars = take(8 * (kSR * NS + 2 * kSNE + kSR));
    s.mail = take(16 * 2 * W);
    s.red = take(16 * W + 16);
    s.tscr = take(bwd ? 16 * 32 : 0);            // 16 scratch bytes per lane: where lanes without a group put their occupancies
    s.tgt = take(4 * round_up(Sp + 1, 128));
    s.em = take(4 * kSNE * star2_em_floats(NLmax));
    s.st = take(bwd ? 4 * kSR * 12 * NLmax : 0);          // stored rows of the other side: one slot per row warp
    s.rows = take(4 * kSR * NS * V);
    s.total = o;
    return s;
}

// s_off word of a position: byte offset of its class in a row (label * 4) | 1 "a lower position holds the same class"
// | 2 "no label state" (position L, the final star)
constexpr int kS2NotFirst = 1, kS2NoLabel = 2;

// Row statistics: the non-blank sum s_nb = sum_{c >= 1} 2^(x_c log2e - m2), the blank term e0, the shift m2 (0 when the
// plain sums stay inside the fp32 range).  V % 4 == 0.
struct StarStats { float s_nb, e0, m2; };
__device__ __forceinline__ StarStats star_row_stats(const float* row, int V4, int lane) {
    const float4* r4 = (const float4*)row;
    StarStats st;
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 4
    for (int c = lane; c < V4; c += 32) {
        const float4 v = r4[c];
        const float ex = ex2f(v.x * kLog2e);
        s0 += ((c == 0) ? 0.0f : ex) + ex2f(v.y * kLog2e);
        s1 += ex2f(v.z * kLog2e) + ex2f(v.w * kLog2e);
    }
    st.s_nb = warp_sum(s0 + s1);
    st.e0 = ex2f(row[0] * kLog2e);
    st.m2 = 0.0f;
    const float tot = st.s_nb + st.e0;
    if (tot > 1e-30f && tot < 1e30f) return st;
    float mx = -CUDART_INF_F;
    for (int c = lane; c < V4; c += 32) {
        const float4 v = r4[c];
        mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    mx = warp_max(mx);
    const float m2 = mx * kLog2e;
    s0 = 0.0f; s1 = 0.0f;
    for (int c = lane; c < V4; c += 32) {
        const float4 v = r4[c];
        const float ex = ex2f(fmaf(v.x, kLog2e, -m2));
        s0 += ((c == 0) ? 0.0f : ex) + ex2f(fmaf(v.y, kLog2e, -m2));
        s1 += ex2f(fmaf(v.z, kLog2e, -m2)) + ex2f(fmaf(v.w, kLog2e, -m2));
    }
    st.s_nb = warp_sum(s0 + s1);
    st.e0 = ex2f(fmaf(row[0], kLog2e, -m2));
    st.m2 = m2;
    return st;
}
// {l2, s_nb, m2, pscale}: l2 = log2 sum_c 2^(x_c log2e) (0 at the log-prob boundary), pscale = 2^(m2 - l2)
__device__ __forceinline__ float4 star_row_norms(StarStats st, int from_logits) {
    const float tot = st.s_nb + st.e0;
    const float l2 = from_logits ? st.m2 + log2f(tot) : 0.0f;
    const float pscale = from_logits ? __frcp_rn(tot) : ex2f(st.m2);
    return make_float4(l2, st.s_nb, st.m2, pscale);
}

// Blank, label and star emissions of one row into a slot: four independent positions per lane and iteration.
__device__ __forceinline__ void star_gather_row(float* em, int NA, const float* row, const int* s_off, int L, int lane, float4 nm) {
    const QRowNorm rn = s2_row_norm(nm.x);
    if (lane == 0) em[0] = s2_emission(row[0], rn);
    for (int k0 = lane; k0 <= L; k0 += 128) {
        int wd[4]; float xv[4], pl[4], ps[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wd[q] = s_off[k0 + 32 * q];
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = *(const float*)((const char*)row + (wd[q] & ~3));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pl[q] = (wd[q] & kS2NoLabel) ? 0.0f : s2_emission(xv[q], rn);
            ps[q] = s2_star_emission(xv[q], (wd[q] & ~3) != 0, nm.y, nm.z, nm.w);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k0 + 32 * q <= L) { em[8 + k0 + 32 * q] = pl[q]; em[8 + NA + k0 + 32 * q] = ps[q]; }
    }
}

// J consecutive words of a slot / stored row / boundary (J = 4, 2, 1: one 16-, 8- or 4-byte access)
template <int J, typename T> __device__ __forceinline__ void ldv(const void* p, T (&v)[J]) {
    if constexpr (J == 4) { const int4 t = *(const int4*)p; const int w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = std::is_same<T, float>::value ? (T)__int_as_float(w[c]) : (T)w[c]; }
    else if constexpr (J == 2) { const int2 t = *(const int2*)p; const int w[2] = {t.x, t.y};
#pragma unroll
        for (int c = 0; c < 2; ++c) v[c] = std::is_same<T, float>::value ? (T)__int_as_float(w[c]) : (T)w[c]; }
    else { const int w = *(const int*)p; v[0] = std::is_same<T, float>::value ? (T)__int_as_float(w) : (T)w; }
}
template <int J> __device__ __forceinline__ int wbits(float x) { return __float_as_int(x); }
template <int J> __device__ __forceinline__ int wbits(int x) { return x; }
template <int J, typename T> __device__ __forceinline__ void stv(void* p, const T (&v)[J]) {
    if constexpr (J == 4) *(int4*)p = make_int4(wbits<J>(v[0]), wbits<J>(v[1]), wbits<J>(v[2]), wbits<J>(v[3]));
    else if constexpr (J == 2) *(int2*)p = make_int2(wbits<J>(v[0]), wbits<J>(v[1]));
    else *(int*)p = wbits<J>(v[0]);
}

// label indices a = position + 4 of an utterance (positions -4 .. L), padded to a multiple of 4
__host__ __device__ inline int star2_na(int L) { return (L + 8) & ~3; }

struct QCfg {
    int g;             // my position group (clamped into the slot): positions J g - 4 .. J g + J - 5
    unsigned allowed;  // skip-transition bits of my J quads
    bool live;         // the lane owns a group of the stored rows
};
template <int J>
__device__ __forceinline__ QCfg star_lane_cfg(int gl, int dir, int L, int NL, int NLmax, const int* y) {
    QCfg cf;
    cf.live = gl < NL;
    const int g = dir ? NL - 1 - gl : gl;
    cf.g = min(max(g, 0), 4 * NLmax / J - 1);
    cf.allowed = s2_allowed<J>(g, dir, L, y, kLabelMask);
    return cf;
}
// the virtual source: mass 1 on the label below the first quad (alpha) / above the final quad (beta)
template <int J>
__device__ __forceinline__ void star_inject(QLane<J>& s, int gl, int dir, int L, int NL) {
    const int a_inj = dir ? L + 4 : 3;
    const int gi = a_inj / J, ci = a_inj % J;
    if (gl == (dir ? NL - 1 - gi : gi)) {
#pragma unroll
        for (int c = 0; c < J; ++c)
            if (c == ci) { s.lb[c] = 1.0f; s.e[c] = 0; }
    }
}

// neighbour values entering my lane: from the previous lane of the side, or the mailbox the warp below filled last step
template <int DIR, int J>
__device__ __forceinline__ void star_fetch(const QLane<J>& s, int w, int lane, const int4* mail_prev, float& n0, float& nl, int& ne) {
    constexpr int C = DIR ? 0 : J - 1;
    nl = __shfl_up_sync(0xffffffffu, s.lb[C], 1);
    ne = __shfl_up_sync(0xffffffffu, s.e[C], 1);
    n0 = DIR ? __shfl_up_sync(0xffffffffu, s.b0[C], 1) : 0.0f;
    if (lane == 0) {
        int4 in = make_int4(0, 0, kQVoidE, 0);
        if (w > 0) in = mail_prev[w - 1];
        n0 = __int_as_float(in.x); nl = __int_as_float(in.y); ne = in.z;
    }
}
template <int DIR, int J>
__device__ __forceinline__ int4 star_mail(const QLane<J>& s) {
    constexpr int C = DIR ? 0 : J - 1;
    return make_int4(__float_as_int(s.b0[C]), __float_as_int(s.lb[C]), s.e[C], 0);
}

// ------------------------------------------------------------------------------------ forward ---
// grid 2N (CTA c: utterance order[c / 2], direction c % 2), block 32 (W + R).
template <int W, int R, int J, int MINB>
__global__ void __launch_bounds__(32 * (W + R), MINB) star2_fwd_kernel(Star2Params p) {
    constexpr int kSR = R, kSNE = 2 * R;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.order[blockIdx.x >> 1], dir = blockIdx.x & 1;
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.hdr[0] = p.pen;
    // feasible iff there is a frame per label plus one between equal neighbours (labels have no self loop)
    if (mt.z || Tn == 0 || Tn < L + mt.w) {
        if (dir == 0 && threadIdx.x == 0) {
            const float v = mt.z ? CUDART_NAN_F : CUDART_INF_F;
            p.loss[n] = v; p.loss_ws[n] = v;
        }
        return;
    }
    const int NA_n = star2_na(L), NL = NA_n / J;   // label indices and lanes (groups of J quads) of this utterance
    const int Wn = (NL + 31) >> 5;                 // trellis warps this utterance needs
    const int NS = p.NS, V = p.V, EMF = p.EMF, NA = p.NA;
    const Star2Smem sm = star2_smem(W, R, NS, V, p.Sp, p.NLmax, false);
    uint64_t* row_full = (uint64_t*)(smem + sm.bars);            // [kSR][NS]
    uint64_t* em_full = row_full + kSR * NS;                     // [kSNE]
    uint64_t* em_empty = em_full + kSNE;                         // [kSNE]
    int4* mail = (int4*)(smem + sm.mail);                        // [2][W]
    double* redd = (double*)(smem + sm.red);                     // [W]
    int* redi = (int*)(redd + W);                                // [W] + flag
    int* s_off = (int*)(smem + sm.tgt);
    float* s_em = (float*)(smem + sm.em);
    float* s_rows = (float*)(smem + sm.rows);
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;
    const int Ks = min(L + 1, p.S);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSR * NS; ++i) mbar_init(&row_full[i], 1);
        for (int i = 0; i < kSNE; ++i) { mbar_init(&em_full[i], 32); mbar_init(&em_empty[i], 32 * Wn); }
    }
    for (int i = threadIdx.x; i < kSNE * EMF; i += blockDim.x) s_em[i] = 0.0f;
    for (int k = threadIdx.x; k < round_up(L + 1, 128); k += blockDim.x) {
        int wd = kS2NotFirst | kS2NoLabel;
        if (k < Ks) {
            const int t = p.tgt[(size_t)n * p.Sp + k];
            wd = ((t & kLabelMask) << 2) | ((t & kNotFirst) ? kS2NotFirst : 0) | (k >= L ? kS2NoLabel : 0);
        }
        s_off[k] = wd;
    }
    mbar_init_fence();
    __syncthreads();

    if (warp >= W) {
        // ---------------------------------------------------------------------- row warps ---
        const int r = warp - W;
        float* wrows = s_rows + (size_t)r * NS * V;
        uint64_t* wbar = row_full + r * NS;
        const float* xb = p.x + (long long)n * p.sx_n;
        const int nrows = (steps1 > r) ? (steps1 - 1 - r) / kSR + 1 : 0;
        const int V4 = V >> 2;
        // stage of row k = k % NS (running counters: NS need not be a power of two)
        auto issue = [&](int k, int stg) {
            const int i = r + k * kSR, t = dir ? Tn - 1 - i : i;
            if (lane == 0) {
                mbar_expect_tx(&wbar[stg], (uint32_t)V * 4u);
                bulk_g2s(wrows + stg * V, xb + (long long)t * p.sx_t, (uint32_t)V * 4u, &wbar[stg]);
            }
        };
        for (int k = 0; k < min(NS, nrows); ++k) issue(k, k);
        int stg = 0; uint32_t par = 0;
        for (int k = 0; k < nrows; ++k) {
            const int i = r + k * kSR, t = dir ? Tn - 1 - i : i;
            const float* row = wrows + stg * V;
            mbar_wait(&wbar[stg], par);
            const float4 nm = star_row_norms(star_row_stats(row, V4, lane), p.from_logits);
            const int slot = i & (kSNE - 1), use = i / kSNE;
            if (use > 0) mbar_wait_sleep(&em_empty[slot], (uint32_t)(use - 1) & 1u, 128);
            star_gather_row(s_em + slot * EMF, NA, row, s_off, L, lane, nm);
            mbar_arrive(&em_full[slot]);
            if (lane == 0) p.stat[(size_t)n * p.T + t] = nm;
            __syncwarp();
            if (k + NS < nrows) issue(k + NS, stg);
            if (++stg == NS) { stg = 0; par ^= 1u; }
        }
        return;
    }
    if (warp >= Wn) return;

    // ------------------------------------------------------------------------ trellis warps ---
    const int w = warp;
    const int nthr = 32 * Wn;
    const int gl = 32 * w + lane;
    const QCfg cfg = star_lane_cfg<J>(gl, dir, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
    QLane<J> s;
    s2_lane_clear(s);
    star_inject(s, gl, dir, L, NL);
    if (lane == 31) mail[1 * W + w] = dir ? star_mail<1>(s) : star_mail<0>(s);
    side_barrier(nthr);
    const float* emp = s_em + 4 + J * cfg.g;                // my J label emissions in slot 0
    const float pen = p.pen;

    auto sweep = [&](auto dirc) {
        constexpr int DIR = decltype(dirc)::value;
        // lanes without a group of their own store to 64 scratch bytes of their own instead of branching around the stores
        // (a divergent branch needs a convergence barrier: see s2_quad_sums)
        int* trow = cfg.live ? p.tr + ((size_t)n * p.T + (DIR ? Tn - 1 : 0)) * p.SPL + J * cfg.g
                             : p.scr + (((size_t)n * 2 + DIR) * 32 + lane) * 16;
        const int NL4 = cfg.live ? NA_n : 4;
        const long long tstep = !cfg.live ? 0 : (DIR ? -(long long)p.SPL : (long long)p.SPL);
        S2P_DECL(4);
        // The slot of step i + 1 is tested (non-blocking) before the arithmetic of step i and, when it is there, read
        // before the step's barrier: the latency of the phase test and of the loads is off the step's critical path.
        float pb = 0.0f, pl[J] = {}, ps[J] = {};
        auto load_slot = [&](int slot, float& b, float (&l)[J], float (&sx)[J]) {
            b = s_em[slot * EMF];
            ldv<J>(emp + slot * EMF, l);
            ldv<J>(emp + NA + slot * EMF, sx);
        };
        if (steps1 > 0) {
            mbar_wait_sleep(&em_full[0], 0u, 32);
            load_slot(0, pb, pl, ps);
            mbar_arrive(&em_empty[0]);
        }
        for (int i = 0; i < steps1; ++i) {
            const int i1 = i + 1, slot1 = i1 & (kSNE - 1);
            const uint32_t par1 = (uint32_t)(i1 / kSNE) & 1u;
            const bool more = i1 < steps1;
            S2P_MARK(3);
            const bool rdy = more & mbar_test(&em_full[slot1], par1);
            S2P_MARK(0);
            float n0, nl; int ne;
            star_fetch<DIR>(s, w, lane, mail + ((i + 1) & 1) * W, n0, nl, ne);
            QSums<J> q;
            s2_quad_sums<DIR>(s, cfg.allowed, n0, nl, ne, q);
            s2_quad_emit(s, q, pb, pl, ps, pen);
            if (lane == 31) mail[(i & 1) * W + w] = star_mail<DIR>(s);
            float npb, npl[J], nps[J];          // (assigned on both ways to their only use, behind `more`)
            if (rdy) {        // warp-uniform; (before this step's global stores: the release of the arrival would wait for them)
                load_slot(slot1, npb, npl, nps);
                mbar_arrive(&em_empty[slot1]);
            }
            stv<J>(trow, s.lb);
            stv<J>(trow + NL4, s.st);
            stv<J>(trow + 2 * NL4, s.e);
            trow += tstep;
            S2P_MARK(1);
            side_barrier(nthr);
            S2P_MARK(2);
            if (more) {
                if (!rdy) {
                    mbar_wait_sleep(&em_full[slot1], par1, 32);
                    load_slot(slot1, npb, npl, nps);
                    mbar_arrive(&em_empty[slot1]);
                }
                pb = npb;
#pragma unroll
                for (int c = 0; c < J; ++c) { pl[c] = npl[c]; ps[c] = nps[c]; }
            }
        }
        if (w == 0) S2P_DUMP(16 + 8 * (int)blockIdx.x, 4, steps1);      // [em wait, compute, barrier, loop]
    };
    if (dir) sweep(std::integral_constant<int, 1>{}); else sweep(std::integral_constant<int, 0>{});

    // ---- the meeting: leave my boundary state; whoever arrives second forms Z from both ----
    int* mybound = p.bound + ((size_t)n * 2 + dir) * p.BW;
    const int NL4 = NA_n;
    if (cfg.live) {
        stv<J>(mybound + J * cfg.g, s.b0);
        stv<J>(mybound + NL4 + J * cfg.g, s.st);
        stv<J>(mybound + 2 * NL4 + J * cfg.g, s.b1);
        stv<J>(mybound + 3 * NL4 + J * cfg.g, s.lb);
        stv<J>(mybound + 4 * NL4 + J * cfg.g, s.e);
    }
    __threadfence();
    side_barrier(nthr);
    if (threadIdx.x == 0) { redi[W] = atomicAdd(&p.cnt[n], 1); __threadfence(); }
    side_barrier(nthr);
    if (redi[W] == 0) return;                                  // the other side is still sweeping: it will do it
    {
        // Z = sum over the states of (alpha's pre-emission sums of the meeting frame) x (beta's boundary).  Whichever
        // side gets here runs the SAME arithmetic on the two stored boundaries (bit-identical repeats).
        const int* ba = p.bound + ((size_t)n * 2 + 0) * p.BW;
        const int* ob = p.bound + ((size_t)n * 2 + 1) * p.BW;
        const QCfg ca = star_lane_cfg<J>(gl, 0, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
        QLane<J> sa;
        float bo[4][J]; int be[J];
#pragma unroll
        for (int c = 0; c < J; ++c) {
            const int a = J * ca.g + c;
            sa.b0[c] = ca.live ? __int_as_float(__ldcg(ba + a)) : 0.0f;
            sa.st[c] = ca.live ? __int_as_float(__ldcg(ba + NL4 + a)) : 0.0f;
            sa.b1[c] = ca.live ? __int_as_float(__ldcg(ba + 2 * NL4 + a)) : 0.0f;
            sa.lb[c] = ca.live ? __int_as_float(__ldcg(ba + 3 * NL4 + a)) : 0.0f;
            sa.e[c] = ca.live ? __ldcg(ba + 4 * NL4 + a) : kQVoidE;
#pragma unroll
            for (int j = 0; j < 4; ++j) bo[j][c] = ca.live ? __int_as_float(__ldcg(ob + j * NL4 + a)) : 0.0f;
            be[c] = ca.live ? __ldcg(ob + 4 * NL4 + a) : kQVoidE;
        }
        if (lane == 31) mail[w] = star_mail<0>(sa);
        side_barrier(nthr);
        float n0, nl; int ne;
        star_fetch<0>(sa, w, lane, mail, n0, nl, ne);
        QSums<J> q;
        s2_quad_sums<0>(sa, ca.allowed, n0, nl, ne, q);
        float zm[4 * J]; int zx[4 * J];
        int pm = 4 * kQVoidE;
#pragma unroll
        for (int c = 0; c < J; ++c) {
            const float m[4] = {q.w0[c] * bo[0][c], q.vs[c] * bo[1][c], q.u1[c] * bo[2][c], q.vl[c] * bo[3][c]};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                zm[4 * c + j] = m[j];
                zx[4 * c + j] = (m[j] > 0.0f) ? sa.e[c] + be[c] + (__float_as_int(m[j]) >> 23) : 4 * kQVoidE;
                pm = max(pm, zx[4 * c + j]);
            }
        }
        // pm: the largest (scale + biased fp32 exponent) of any term; terms are summed relative to it in float64
        pm = __reduce_max_sync(0xffffffffu, pm);
        if (lane == 0) redi[w] = pm;
        side_barrier(nthr);
        for (int x = 0; x < Wn; ++x) pm = max(pm, redi[x]);
        const bool feasible = pm > kVoidETest;
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 4 * J; ++j) {
            if (zm[j] > 0.0f) {
                const float mant = __int_as_float((__float_as_int(zm[j]) & 0x007fffff) | 0x3f800000);   // in [1, 2)
                const int rel = zx[j] - pm;                                      // <= 0
                if (rel > -1000) sum += scalbn((double)mant, rel);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) redd[w] = sum;
        side_barrier(nthr);
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int x = 0; x < Wn; ++x) tot += redd[x];
            // Z = tot * 2^(pm - 127): every term is mant * 2^(its scale + biased exponent - 127)
            float v = CUDART_INF_F;
            int4 zi = make_int4(1 << 29, __float_as_int(1.0f), 0, 0);
            if (feasible && tot > 0.0) {
                const int ex = ilogb(tot);
                const double log2z = (double)(pm - 127) + log2(tot);
                v = (float)(-log2z * kLn2);
                zi.x = pm - 127 + ex;
                zi.y = __float_as_int((float)(1.0 / scalbn(tot, -ex)));
            }
            p.loss[n] = v; p.loss_ws[n] = v;
            p.zinfo[n] = zi;
        }
    }
}

// ----------------------------------------------------------------------------------- backward ---
// grid 2N, block 32 (W + R).  d loss / d x[c] = gout (p_c (delta - G + H_c) - occ_c): delta = 1 through the fused
// log-softmax and 0 at the log-prob boundary, G = sum_k h_k, H_c = the sum of h_k over the stars that exclude c,
// occ_c the label occupancy of class c (the blank: one minus the label and star occupancies).  Rows t >= T_n and every
// row of an infeasible or invalid utterance are zero.
template <int W, int R, int J, int MINB>
__global__ void __launch_bounds__(32 * (W + R), MINB) star2_bwd_kernel(Star2Params p) {
    constexpr int kSR = R, kSNE = 2 * R;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.order[blockIdx.x >> 1], dir = blockIdx.x & 1;
    const int4 mt = p.meta[n];
    const int L = mt.y;
    const float lossn = p.loss_ws[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x;      // NaN / inf loss: all-zero gradient
    const int NS = p.NS, V = p.V, EMF = p.EMF, NA = p.NA, V4 = V >> 2;
    float* gb = p.gx + (long long)n * p.sg_n;
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;
    const int nsteps2 = Tn - steps1;

    if (warp >= W) {
        // rows past the end of the utterance: zero, shared between the two sides' row warps
        const int r = warp - W;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = Tn + dir + 2 * r; t < p.T; t += 2 * kSR) {
            float4* dst = (float4*)(gb + (long long)t * p.sg_t);
            for (int c = lane; c < V4; c += 32) dst[c] = z;
        }
    }
    if (nsteps2 <= 0) return;

    const int NA_n = star2_na(L), NL = NA_n / J;
    const int Wn = (NL + 31) >> 5;
    const int Ks = min(L + 1, p.S);
    const Star2Smem sm = star2_smem(W, R, NS, V, p.Sp, p.NLmax, true);
    uint64_t* row_full = (uint64_t*)(smem + sm.bars);            // [kSR][NS]
    uint64_t* em_full = row_full + kSR * NS;                     // [kSNE]
    uint64_t* occ_full = em_full + kSNE;                         // [kSNE]
    uint64_t* st_full = occ_full + kSNE;                         // [kSR]
    int4* mail = (int4*)(smem + sm.mail);
    int* s_off = (int*)(smem + sm.tgt);
    const int* nfl = p.nflist + (size_t)n * p.NF;                // later occurrences of a class (ctc2_prep_kernel)
    float* s_em = (float*)(smem + sm.em);
    int* s_st = (int*)(smem + sm.st);
    float* s_rows = (float*)(smem + sm.rows);
    const float g = p.gout[n];
    const float delta = p.from_logits ? 1.0f : 0.0f;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSR * NS; ++i) mbar_init(&row_full[i], 1);
        for (int i = 0; i < kSNE; ++i) { mbar_init(&em_full[i], 32); mbar_init(&occ_full[i], 32 * Wn); }
        for (int i = 0; i < kSR; ++i) mbar_init(&st_full[i], 1);
    }
    for (int i = threadIdx.x; i < kSNE * EMF; i += blockDim.x) s_em[i] = 0.0f;
    const int2 nf = p.nfhdr[n];                                  // {entries in rank groups of 32, serial tail}
    for (int k = threadIdx.x; k < round_up(L + 1, 128); k += blockDim.x) {
        int wd = kS2NotFirst | kS2NoLabel;
        if (k < Ks) {
            const int t = p.tgt[(size_t)n * p.Sp + k];
            wd = ((t & kLabelMask) << 2) | ((t & kNotFirst) ? kS2NotFirst : 0) | (k >= L ? kS2NoLabel : 0);
        }
        s_off[k] = wd;
    }
    mbar_init_fence();
    __syncthreads();

    if (warp >= W) {
        // ---------------------------------------------------------------------- row warps ---
        const int r = warp - W;
        float* wrows = s_rows + (size_t)r * NS * V;
        uint64_t* wbar = row_full + r * NS;
        const float* xb = p.x + (long long)n * p.sx_n;
        const int nrows = (nsteps2 > r) ? (nsteps2 - 1 - r) / kSR + 1 : 0;
        auto frame = [&](int k) { const int i = steps1 + r + k * kSR; return dir ? Tn - 1 - i : i; };
        S2P_DECL(6);
        // stage of row k = k % NS
        auto issue = [&](int k, int stg) {
            if (lane == 0) {
                mbar_expect_tx(&wbar[stg], (uint32_t)V * 4u);
                bulk_g2s(wrows + stg * V, xb + (long long)frame(k) * p.sx_t, (uint32_t)V * 4u, &wbar[stg]);
            }
        };
        // the row the OTHER side stored for the frame of my row k, into my stored-row slot: issued once the trellis warps
        // have read the slot for my row k - 1 (I have seen their occupancies), R steps before they need it
        const uint32_t st_bytes = (uint32_t)(3 * NA_n) * 4u;
        const int* tr_n = p.tr + (size_t)n * p.T * p.SPL;
        auto issue_st = [&](int k) {
            if (lane == 0) {
                mbar_expect_tx(&st_full[r], st_bytes);
                bulk_g2s(s_st + r * p.SPL, tr_n + (size_t)frame(k) * p.SPL, st_bytes, &st_full[r]);
            }
        };
        if (nrows > 0) issue_st(0);
        float4 nm_next = make_float4(0.f, 0.f, 0.f, 0.f);       // row statistics, fetched one row ahead
        if (nrows > 0) nm_next = p.stat[(size_t)n * p.T + frame(0)];
        // later occurrences of a class: the first four rank groups live in registers (a global load per group and row
        // would sit on every row's critical path)
        int nfe[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) nfe[j] = (32 * j < nf.x) ? __ldg(nfl + 32 * j + lane) : -1;
        // pre(k): gather the emissions of row k for the trellis warps, then turn the row into g * p_c in place.
        auto pre = [&](int k, int stg, uint32_t par) {
            const int i2 = r + k * kSR, t = frame(k);
            float* row = wrows + stg * V;
            const float4 nm = nm_next;
            if (k + 1 < nrows) nm_next = p.stat[(size_t)n * p.T + frame(k + 1)];
            S2P_MARK(5);
            mbar_wait(&wbar[stg], par);
            S2P_MARK(0);
            // (slot i2 % kSNE was last used by my own row k - 2, whose occupancies I consumed in program order)
            star_gather_row(s_em + (i2 & (kSNE - 1)) * EMF, NA, row, s_off, L, lane, nm);
            mbar_arrive(&em_full[i2 & (kSNE - 1)]);
            __syncwarp();
            float4* r4 = (float4*)row;
            const float l2 = nm.x;
#pragma unroll 4
            for (int c = lane; c < V4; c += 32) {
                float4 v = r4[c];
                v.x = g * ex2f(fmaf(v.x, kLog2e, -l2)); v.y = g * ex2f(fmaf(v.y, kLog2e, -l2));
                v.z = g * ex2f(fmaf(v.z, kLog2e, -l2)); v.w = g * ex2f(fmaf(v.w, kLog2e, -l2));
                r4[c] = v;
            }
            __syncwarp();
            S2P_MARK(1);
        };
        // post(k): the occupancies the trellis warps left in the slot -> the dense gradient row, stored.
        auto post = [&](int k, int stg) {
            const int i2 = r + k * kSR, t = frame(k);
            float* row = wrows + stg * V;
            float* em = s_em + (i2 & (kSNE - 1)) * EMF;
            S2P_MARK(5);
            mbar_wait_sleep(&occ_full[i2 & (kSNE - 1)], (uint32_t)(i2 / kSNE) & 1u, 128);
            if (k + 1 < nrows) issue_st(k + 1);
            S2P_MARK(2);
            // per position: a_k = g p_{y_k} h_k - g gamma(label k), parked in the label-occupancy word; bs = label + star
            // occupancies, G = sum of h
            float bs = 0.0f, G = 0.0f;
            for (int k0 = lane; k0 <= L; k0 += 128) {
                int wd[4]; float gl[4], gs[4], h[4], rv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = k0 + 32 * q <= L;
                    wd[q] = s_off[k0 + 32 * q];
                    gl[q] = in ? em[8 + k0 + 32 * q] : 0.0f;
                    h[q] = in ? em[8 + NA + k0 + 32 * q] : 0.0f;
                    gs[q] = in ? em[8 + 2 * NA + k0 + 32 * q] : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) rv[q] = *(const float*)((const char*)row + (wd[q] & ~3));
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    bs += gl[q] + gs[q]; G += h[q];
                    const float a = (((wd[q] & ~3) != 0) ? rv[q] * h[q] : 0.0f) - g * gl[q];
                    if (k0 + 32 * q <= L) em[8 + k0 + 32 * q] = a;
                }
            }
            bs = warp_sum(bs); G = warp_sum(G);
            const float r0 = row[0];
            __syncwarp();
            float4* r4 = (float4*)row;
            const float sc = delta - G;
#pragma unroll 4
            for (int c = lane; c < V4; c += 32) {
                float4 v = r4[c];
                v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                r4[c] = v;
            }
            __syncwarp();
            if (lane == 0) row[0] = r0 * delta - g * (1.0f - bs);
            __syncwarp();
            // first occurrences of a class: all distinct, four per lane and iteration
            for (int k0 = lane; k0 <= L; k0 += 128) {
                int wd[4]; float a[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    wd[q] = s_off[k0 + 32 * q];
                    a[q] = (k0 + 32 * q <= L) ? em[8 + k0 + 32 * q] : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (!(wd[q] & kS2NotFirst)) *(float*)((char*)row + (wd[q] & ~3)) += a[q];
            }
            __syncwarp();
            // later occurrences: by occurrence rank, one rank (distinct classes) per 32 entries: deterministic, no atomics
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (32 * j < nf.x) {
                    if (nfe[j] >= 0) row[nfe[j] >> 10] += em[8 + (nfe[j] & 1023)];
                    __syncwarp();
                }
            }
            for (int e0 = 128; e0 < nf.x; e0 += 32) {
                const int e = __ldg(nfl + e0 + lane);
                if (e >= 0) row[e >> 10] += em[8 + (e & 1023)];
                __syncwarp();
            }
            if (nf.y > 0 && lane == 0)
                for (int x = nf.x; x < nf.x + nf.y; ++x) { const int e = __ldg(nfl + x); row[e >> 10] += em[8 + (e & 1023)]; }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(gb + (long long)t * p.sg_t, row, (uint32_t)V * 4u); bulk_commit(); }
            S2P_MARK(3);
        };
        auto next = [&](int& stg, uint32_t& par) { if (++stg == NS) { stg = 0; par ^= 1u; } };
        if (NS >= 3) {
            // NS - 1 rows in flight ahead of the one being finished: while row k waits for its occupancies and is stored,
            // row k + 1 is gathered (its emissions are what the trellis warps wait for) and rows k + 2 .. k + NS - 1 load.
            // Row k + NS - 1 goes into the stage of row k - 1, once that row's store has read it.
            for (int k = 0; k < min(NS - 1, nrows); ++k) issue(k, k);
            int s_post = 0, s_pre = 0, s_ld = (NS - 1) % NS; uint32_t p_pre = 0, p_unused = 0;
            if (nrows > 0) { pre(0, s_pre, p_pre); next(s_pre, p_pre); }
            for (int k = 0; k < nrows; ++k) {
                if (k + 1 < nrows) { pre(k + 1, s_pre, p_pre); next(s_pre, p_pre); }
                if (k + NS - 1 < nrows) {
                    S2P_MARK(5);
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    S2P_MARK(4);
                    issue(k + NS - 1, s_ld);
                    next(s_ld, p_unused);
                }
                post(k, s_post);
                next(s_post, p_unused);
            }
        } else if (NS == 2) {
            for (int k = 0; k < min(2, nrows); ++k) issue(k, k);
            if (nrows > 0) pre(0, 0, 0);
            for (int k = 0; k < nrows; ++k) {
                if (k + 1 < nrows) pre(k + 1, (k + 1) & 1, (uint32_t)((k + 1) >> 1) & 1u);
                post(k, k & 1);
                if (k + 2 < nrows) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    issue(k + 2, k & 1);
                }
            }
        } else {
            for (int k = 0; k < nrows; ++k) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                issue(k, 0);
                pre(k, 0, (uint32_t)k & 1u);
                post(k, 0);
            }
        }
        if (r == 0) S2P_DUMP(48 + 8 * (int)blockIdx.x, 6, nrows);     // [row wait, pre, occ wait, post, store-read wait, other]
        if (lane == 0) bulk_wait_all<0>();
        return;
    }
    if (warp >= Wn) return;

    // ------------------------------------------------------------------------ trellis warps ---
    const int w = warp;
    const int nthr = 32 * Wn;
    const int gl = 32 * w + lane;
    const QCfg cfg = star_lane_cfg<J>(gl, dir, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
    const int NL4 = NA_n;
    QLane<J> s;
    s2_lane_clear(s);
    if (cfg.live) {
        const int* mybound = p.bound + ((size_t)n * 2 + dir) * p.BW + J * cfg.g;
        ldv<J>(mybound, s.b0); ldv<J>(mybound + NL4, s.st); ldv<J>(mybound + 2 * NL4, s.b1); ldv<J>(mybound + 3 * NL4, s.lb);
        ldv<J>(mybound + 4 * NL4, s.e);
    }
    const int4 zi = p.zinfo[n];
    const int eZ = zi.x;
    const float rZ = __int_as_float(zi.y);
    const float pen = p.hdr[0];
    if (lane == 31) mail[((steps1 + 1) & 1) * W + w] = dir ? star_mail<1>(s) : star_mail<0>(s);
    float* emp = s_em + 4 + J * cfg.g;                      // my J label emissions / occupancies in slot 0
    float* s_scratch = (float*)(smem + sm.tscr) + 4 * lane;
    const int* stp = s_st + J * cfg.g;                      // the other side's label states of my group in slot 0
    side_barrier(nthr);

    auto sweep = [&](auto dirc) {
        constexpr int DIR = decltype(dirc)::value;
        S2P_DECL(5);
        // software-pipelined slot hand-over as in the forward kernel; a step's inputs are its emission slot (i2 % 2R)
        // and the stored-row slot of the row warp that owns the frame (i2 % R)
        float pb = 0.0f, pl[J] = {}, ps[J] = {}, lbo[J] = {}, sto[J] = {};
        int eo[J];
#pragma unroll
        for (int c = 0; c < J; ++c) eo[c] = kQVoidE;
        auto load_slot = [&](int slot, int ss, float& b, float (&l)[J], float (&sx)[J], float (&oo)[J], float (&so)[J], int (&ee)[J]) {
            b = s_em[slot * EMF];
            ldv<J>(emp + slot * EMF, l);
            ldv<J>(emp + NA + slot * EMF, sx);
            // (lanes without a group of their own read whatever lies at their clamped position: their occupancies go to
            // the scratch word)
            ldv<J>(stp + ss * p.SPL, oo); ldv<J>(stp + ss * p.SPL + NL4, so);
            ldv<J>(stp + ss * p.SPL + 2 * NL4, ee);
        };
        if (nsteps2 > 0) {
            mbar_wait_sleep(&em_full[0], 0u, 32);
            mbar_wait(&st_full[0], 0u);
            load_slot(0, 0, pb, pl, ps, lbo, sto, eo);
        }
        for (int i2 = 0; i2 < nsteps2; ++i2) {
            const int i = steps1 + i2;
            const int slot = i2 & (kSNE - 1);
            const int j1 = i2 + 1, slot1 = j1 & (kSNE - 1), ss1 = j1 & (kSR - 1);
            const uint32_t par1 = (uint32_t)(j1 / kSNE) & 1u, spar1 = (uint32_t)(j1 / kSR) & 1u;
            const bool more = j1 < nsteps2;
            S2P_MARK(4);
            const bool t_em = mbar_test(&em_full[slot1], par1), t_st = mbar_test(&st_full[ss1], spar1);   // both in flight
            const bool rdy = more & t_em & t_st;
            S2P_MARK(0);
            float n0, nl; int ne;
            star_fetch<DIR>(s, w, lane, mail + ((i + 1) & 1) * W, n0, nl, ne);
            QSums<J> q;
            s2_quad_sums<DIR>(s, cfg.allowed, n0, nl, ne, q);
            float ogl[J], ogs[J], oh[J];
            s2_quad_occ(s, q, lbo, sto, eo, eZ, rZ, ps, ogl, ogs, oh);
            s2_quad_emit(s, q, pb, pl, ps, pen);
            if (lane == 31) mail[(i & 1) * W + w] = star_mail<DIR>(s);
            {   // lanes without a group of their own write a scratch word instead of branching around the stores
                float* o0 = cfg.live ? emp + slot * EMF : s_scratch;
                stv<J>(o0, ogl);
                stv<J>(cfg.live ? o0 + NA : s_scratch, oh);
                stv<J>(cfg.live ? o0 + 2 * NA : s_scratch, ogs);
            }
            float npb, npl[J], nps[J], no4[J], ns4[J];      // (assigned on both ways to their only use, behind `more`)
            int ne4[J];
            if (rdy) load_slot(slot1, ss1, npb, npl, nps, no4, ns4, ne4);      // warp-uniform
            mbar_arrive(&occ_full[slot]);
            S2P_MARK(2);
            side_barrier(nthr);
            S2P_MARK(3);
            if (more) {
                if (!rdy) {
#ifdef HAB_STAR2_PROBE
                    s2p_acc[1] += mbar_test(&em_full[slot1], par1) ? 1 : 1000;      // units: waited for the stored row; thousands: for the emissions
#endif
                    mbar_wait_sleep(&em_full[slot1], par1, 32);
                    mbar_wait(&st_full[ss1], spar1);
                    load_slot(slot1, ss1, npb, npl, nps, no4, ns4, ne4);
                }
                pb = npb;
#pragma unroll
                for (int c = 0; c < J; ++c) { pl[c] = npl[c]; ps[c] = nps[c]; lbo[c] = no4[c]; sto[c] = ns4[c]; eo[c] = ne4[c]; }
            }
        }
        if (w == 0) S2P_DUMP(32 + 8 * (int)blockIdx.x, 5, nsteps2);     // [em wait, stored-row wait, compute, barrier, loop]
    };
    if (dir) sweep(std::integral_constant<int, 1>{}); else sweep(std::integral_constant<int, 0>{});
}

}  // namespace hab

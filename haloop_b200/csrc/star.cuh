// star.cuh — star / wildcard CTC (STC) loss + logit gradient for sm_100a.  Replaces
// ha/star.py:65-163 (star_ctc_forward_score), :8-49 (intersperse_stars — the (T,N,2V) tensor is never
// built: only log P_t = log sum_{c>=1} p_c and the per-target log(P_t - p_y) are gathered) and the
// autograd backward.  Same four-kernel shape as ctc.cuh (whose prep kernel is shared).
//
// State layout: position k of the target owns the quad
//     b0 = blank (j=4k)   st = star "anything but y_k" (4k+1)   b1 = blank (4k+2)   lb = label y_k (4k+3)
// and the final quad k = L_n holds (blank, last star, final blank) only.  Transitions, ha/star.py:123-145:
//     b0 <- lb[k-1], b0         st <- b0, st, b1 (+ star_penalty)       b1 <- st, b1
//     lb <- b0, b1, st, and lb[k-1] unless y_k == y_{k-1}               (labels have no self loop)
#pragma once
#include "common.cuh"
#include "ctc.cuh"

namespace hab {

struct StarWs {
    size_t meta, order, tgt, dupnext, loss, lse2, em, tr, total;
    int Sp, E, JWp, SPX;
};

// emission row (32-bit words): [0] float ct (integer row shift)  [1] blank  [2] all-star (log2 P)
// [4+2k] label y_k  [5+2k] star "anything but y_k", each a Q8.24 fixed-point value of log2 p - ct
// (emission_word, ctc.cuh).  The occupancy row written in place (floats): [1] blank occupancy
// [2] G = sum_k h_k  [4+2k] label occupancy  [5+2k] h_k (0 if y_k == 0), h_k = gamma(star k) / (P - p_{y_k}).
__host__ __device__ inline int star_em_floats(int Sp) { return 4 + 2 * Sp; }

__host__ inline StarWs star_ws_layout(int T, int N, int S) {
    StarWs w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.E = star_em_floats(w.Sp);
    w.JWp = round_up((S + 1 + 31) / 32, 4);
    w.SPX = w.JWp + 4 * (S + 1);
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.dupnext = take(sizeof(int) * (size_t)N * w.Sp);
    w.loss = take(sizeof(float) * (size_t)N);
    w.lse2 = take(sizeof(float) * (size_t)N * T);
    w.em = take(sizeof(float) * (size_t)N * T * w.E);
    w.tr = take(sizeof(float) * (size_t)N * T * w.SPX);
    w.total = o;
    return w;
}

// ------------------------------------------------------------------------------------ rows ---
struct StarRowsParams {
    const float* x; long long sx_t, sx_n;
    int T, N, V, S;
    const int4* meta; const int* tgt; int Sp;
    float* lse2; float* em; int E;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

// One warp per (b,t) row: log-softmax statistics, log2 P = log2 sum_{c>=1} p_c, and for every target
// position k < min(L_n + 1, S) the label emission and the star emission log2(P - p_{y_k}) (log2 P when
// y_k == 0); the star in front of position L_n reads targets[n, L_n] (ha/star.py:46-47).  All
// relative to the integer row shift c_t = rint(max(blank, log2 P)).  Row layout: star_em_floats().
template <bool VEC4>
__global__ void __launch_bounds__(kMaxRowWarps * 32) star_rows_kernel(StarRowsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nwarps;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const int Tn = mt.z ? 0 : mt.x, L = mt.y;
    const int Ks = min(L + 1, p.S);
    const int t0 = blockIdx.x * (nw * p.rows_per_warp);
    if (t0 >= Tn) return;
    const int V = p.V, nstage = p.nstage;
    uint64_t* s_bar = (uint64_t*)smem_raw;
    int* s_tgt = (int*)(smem_raw + round_up_sz((size_t)nw * nstage * 8, 128));
    float* s_rows = (float*)((unsigned char*)s_tgt + round_up_sz((size_t)p.Sp * 4, 128));
    for (int k = threadIdx.x; k < Ks; k += blockDim.x) s_tgt[k] = p.tgt[(size_t)n * p.Sp + k] & kLabelMask;
    float* wrows = s_rows + (size_t)warp * nstage * V;
    uint64_t* wbar = s_bar + warp * nstage;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&wbar[s], 1);
    mbar_init_fence();
    __syncthreads();

    int nrows = 0;
    if (t0 + warp < Tn) nrows = min(p.rows_per_warp, (Tn - 1 - t0 - warp) / nw + 1);
    const float* xb = p.x + (long long)n * p.sx_n;
    auto issue = [&](int r) {
        const int stage = r % nstage;
        const float* src = xb + (long long)(t0 + warp + nw * r) * p.sx_t;
        if (p.use_bulk) {
            if (lane == 0) {
                mbar_expect_tx(&wbar[stage], (uint32_t)V * 4u);
                bulk_g2s(wrows + (size_t)stage * V, src, (uint32_t)V * 4u, &wbar[stage]);
            }
        } else {
            for (int c = lane; c < V; c += 32) wrows[(size_t)stage * V + c] = src[c];
        }
    };
    for (int r = 0; r < min(nstage, nrows); ++r) issue(r);

    for (int r = 0; r < nrows; ++r) {
        const int stage = r % nstage;
        if (p.use_bulk) mbar_wait(&wbar[stage], (uint32_t)(r / nstage) & 1u);
        else __syncwarp();
        const float* row = wrows + (size_t)stage * V;
        const int t = t0 + warp + nw * r;
        float mx = -CUDART_INF_F;
        if (VEC4) {
            const float4* r4 = (const float4*)row;
            for (int c = lane; c < (V >> 2); c += 32) {
                float4 v = r4[c];
                mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            }
        } else {
            for (int c = lane; c < V; c += 32) mx = fmaxf(mx, row[c]);
        }
        mx = warp_max(mx);
        const float m2 = mx * kLog2e;
        float s = 0.0f;                                   // sum over the non-blank classes only
        if (VEC4) {
            const float4* r4 = (const float4*)row;
            for (int c = lane; c < (V >> 2); c += 32) {
                float4 v = r4[c];
                const float e0 = ex2f(fmaf(v.x, kLog2e, -m2));
                s += ((c == 0) ? 0.0f : e0) + ex2f(fmaf(v.y, kLog2e, -m2)) +
                     ex2f(fmaf(v.z, kLog2e, -m2)) + ex2f(fmaf(v.w, kLog2e, -m2));
            }
        } else {
            for (int c = lane; c < V; c += 32) s += (c == 0) ? 0.0f : ex2f(fmaf(row[c], kLog2e, -m2));
        }
        s = warp_sum(s);
        const float e0 = ex2f(fmaf(row[0], kLog2e, -m2));
        const float l2 = p.from_logits ? m2 + log2f(s + e0) : 0.0f;
        const float lP = fmaxf(m2 + log2f(s) - l2, kVoid);        // log2 sum_{c>=1} p_c   (ha/star.py:30)
        const float eblank = fmaf(row[0], kLog2e, -l2);
        const float ct = round_int(fmaxf(eblank, lP));            // every other emission is <= log2 P
        // emissions in float-float arithmetic from the fp32 logits, split into int8 integer part + fp32
        // fraction (emission_split, common.cuh).  log2 P and the star terms are (hi, lo) float pairs too.
        const float lgs = log2f(s);
        const float lPh = m2 + lgs;                                        // log2 P + l2, rounded ...
        const float lbb = lPh - m2;
        const float lPl = (m2 - (lPh - lbb)) + (lgs - lbb);                // ... and the rounding error (two-sum)
        float* erow = p.em + ((size_t)n * p.T + t) * p.E;
        // (vh + vl) - l2 - ct  ->  integer part K + fraction f
        auto split2 = [&](float vh, float vl, float& K, float& f) {
            const float s1 = vh - l2;
            const float bb = s1 - vh;
            const float err = (vh - (s1 - bb)) + (-l2 - bb);
            const float Kt = fmaxf(round_int(s1 - ct) + ct, ct - 127.0f);
            K = Kt - ct;
            f = (s1 - Kt) + (err + vl);
        };
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            float Kb, fb, Ka, fa;
            emission_split(row[0], l2, ct, Kb, fb);
            split2(fmaxf(lPh, kVoid), lPl, Ka, fa);
            *(float4*)erow = make_float4(ct, __int_as_float(emission_word(Kb, fb)), __int_as_float(emission_word(Ka, fa)), 0.0f);
        }
        for (int k = lane; k < Ks; k += 32) {
            const int y = s_tgt[k];
            float Kl, fl, Ks2, fs2;
            emission_split(row[y], l2, ct, Kl, fl);
            float add = 0.0f;
            if (y != 0) {
                // logsubexp (ha/star.py:4-5): log2 P + log2(1 - 2^(lab - log2 P)), via expm1 so that a
                // label holding almost all of P does not cancel
                const float d = fminf(fmaf(row[y], kLog2e, -lPh), 0.0f);   // lab - log2 P (l2 cancels)
                add = fmaxf(log2f(-expm1f(d * (float)kLn2)), kVoid);
            }
            // star emission = log2 P + add: fold `add` into the low word (|add| is small unless the label
            // holds nearly all of P, where its own relative error dominates anyway)
            const float sh = lPh + add;
            const float sl = lPl + ((lPh - sh) + add);
            split2(fmaxf(sh, kVoid), sl, Ks2, fs2);
            ((int2*)(erow + 4))[k] = make_int2(emission_word(Kl, fl), emission_word(Ks2, fs2));
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- trellis ---
struct StarTrellisParams {
    int T, N, S;
    const int4* meta; const int* order; const int* tgt; int Sp;
    float* em; int E;          // emission rows in / occupancy rows out, in place (layout: star_em_floats)
    float* tr; int SPX, JWp;   // stored row = [JWp slot bases][4*(L+1) Q11.20 ints: quads]
    float* loss; float* loss_ws;
    float pen2;                // star_penalty in log2 units
    int nstage, G, W;          // ring stages; frames per stage; compute warps per sweep direction
    int dir_bytes;
};

// Same CTA shape and ring protocol as ctc_trellis_kernel (one CTA per utterance, W compute warps +
// one producer warp per sweep side, meet in the middle).  Both sides keep quad k in lane k%32 of slot
// k/32 of warp k/(32 J); beta is not a mirror image of alpha here (labels have no self loop, stars have
// a back edge), so it has its own update and its mailbox runs downwards.
template <int J>
__global__ void __launch_bounds__(448) star_trellis_kernel(StarTrellisParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int W = p.W, G = p.G, nstage = p.nstage;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp >= 2 * W;
    const int dir = producer ? warp - 2 * W : (warp >= W);
    // the beta side takes its warps in reverse order, so each SM sub-partition hosts an alpha warp that is
    // busy early (low label pairs are reached first) next to a beta warp that is busy late
    const int w = producer ? 0 : (dir ? 2 * W - 1 - warp : warp);
    const int n = p.order[blockIdx.x];
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    if (mt.z || Tn == 0) {
        const float v = mt.z ? CUDART_NAN_F : CUDART_INF_F;
        if (threadIdx.x == 0) { p.loss[n] = v; p.loss_ws[n] = v; }
        return;
    }
    const int Q = L + 1, Ks = min(L + 1, p.S);
    const int E = p.E, SPX = p.SPX, JWp = p.JWp, OC = 4 + 2 * p.Sp;
    const int SF_ = trellis_stage_floats(E, SPX, OC, G, W, 2);
    const float Kp = round_int(p.pen2), fp = p.pen2 - Kp;

    unsigned char* db = smem_raw + (size_t)dir * p.dir_bytes;
    float* stages = (float*)db;
    float4* mail = (float4*)(stages + nstage * SF_);              // [2][W]
    double* redm = (double*)(mail + 2 * W);
    float* reds = (float*)(redm + W);
    uint64_t* full = (uint64_t*)(reds + 2 * W);
    if (producer && lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
    __syncthreads();

    float* em_base = p.em + (size_t)n * p.T * E;
    float* tr_base = p.tr + (size_t)n * p.T * SPX;
    const uint32_t occ_bytes = (uint32_t)(4 + round_up(2 * Ks, 4)) * 4u;
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;

    if (producer) {
        trellis_producer<2, false>(stages, SF_, full, nstage, G, W, E, SPX, OC, em_base, tr_base, occ_bytes,
                            Tn, steps1, dir, lane);
        return;
    }

    const bool leader = (w == 0 && lane == 0);
    const int nthr = 32 * W;
    const int barid = 1 + dir;
    const int k0 = 32 * (w * J) + lane;           // my quad in slot 0
    // alpha: may label k be entered from label k-1?   beta: may label k be left for label k+1?
    unsigned allowed = 0, exclude = 0, hasq = 0, hasl = 0, hase = 0;   // exclude: star k excludes a class
    {
        const int* y = p.tgt + (size_t)n * p.Sp;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = k0 + 32 * j;
            if (k < Q) hasq |= 1u << j;
            if (k < L) hasl |= 1u << j;
            if (k < Ks) hase |= 1u << j;
            if (k < Ks && (y[k] & kLabelMask) != 0) exclude |= 1u << j;
            if (dir == 0) {
                if (k < L && (k == 0 || (y[k] & kLabelMask) != (y[k - 1] & kLabelMask))) allowed |= 1u << j;
            } else {
                if (k + 1 < L && (y[k + 1] & kLabelMask) != (y[k] & kLabelMask)) allowed |= 1u << j;
            }
        }
    }
    // my quads cannot be reached before this step (alpha spreads upwards one quad per frame from
    // quad 0, beta downwards from quad L)
    const int first = dir ? L - (32 * (w + 1) * J - 1) - 2 : 32 * (w * J) - 2;
    // ... and none of them can still be on a complete path after this step (one quad per remaining frame)
    const int last = dir ? Tn - 32 * (w * J) + 1 : Tn - L + (32 * (w + 1) * J - 1) + 1;

    SF b0[J], st[J], b1[J], lb[J];
    float base[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { b0[j] = st[j] = b1[j] = lb[j] = sf_void(); base[j] = 0.0f; }

    float IZ = 0.0f, fZ = 0.0f, csum = 0.0f, ct = 0.0f, cin_h = kVoid;
    bool feasible = true;
    float Kb, fb, Kl[J], fl[J], Ks_[J], fs[J];   // blank / label / star emissions of my quad, split; the
                                                  // star's include the penalty paid on entering it
    int s = 0, g = 0, cnt = 0; uint32_t fpar = 0;
    const float* stg = stages;

    auto fetch = [&](int i, int phase_end) {
        if (g == 0) {
            cnt = min(G, phase_end - i);
            stg = stages + s * SF_;
            mbar_wait(&full[s], fpar);
        }
        const int ridx = dir ? cnt - 1 - g : g;
        const float* er = stg + ridx * E;
        const int* ew = (const int*)er;
        ct = er[0];
        csum += ct;
        emission_decode(ew[1], Kb, fb);
        float Ka, fa;
        emission_decode(ew[2], Ka, fa);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = k0 + 32 * j;
            const bool ve = (hase >> j) & 1u, vl = (hasl >> j) & 1u;
            const int2 wv = ve ? ((const int2*)(ew + 4))[k] : make_int2(0, 0);
            float K1, f1, K2, f2;
            emission_decode(wv.x, K1, f1);
            emission_decode(wv.y, K2, f2);
            Kl[j] = vl ? K1 : kVoid;
            fl[j] = vl ? f1 : 0.0f;
            // the star of quad k: its own entry, or the all-star when k == L == S (ha/star.py:47)
            const float ks = ve ? K2 : ((k == L) ? Ka : kVoid);
            const float fs0 = ve ? f2 : ((k == L) ? fa : 0.0f);
            Ks_[j] = fmaxf(ks + Kp, kVoid);
            fs[j] = fs0 + fp;
        }
        return ridx;
    };
    auto advance = [&](int i) {
        if (dir == 0) {
            if (i >= first && i <= last) {
                // previous label, from the lane below (virtual state -1 holds 0.0 before the first frame)
                float ch[J], cl[J];
                {
                    float rh[J], rl[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        rh[j] = __shfl_sync(0xffffffffu, lb[j].h, (lane + 31) & 31);
                        rl[j] = __shfl_sync(0xffffffffu, lb[j].l, (lane + 31) & 31);
                    }
                    float4 in = make_float4((i == 0) ? 0.0f : kVoid, 0.0f, 0.0f, 0.0f);
                    if (w > 0) in = mail[((i - 1) & 1) * W + w - 1];
                    cin_h = in.x;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        ch[j] = lane ? rh[j] : (j ? rh[j ? j - 1 : 0] : in.x);
                        cl[j] = lane ? rl[j] : (j ? rl[j ? j - 1 : 0] : in.y);
                    }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    SF c; c.h = ch[j]; c.l = cl[j];
                    const SF u = lae_sf(st[j], b1[j]);
                    const SF v = lae_sf(u, b0[j]);
                    const SF w0 = lae_sf(c, b0[j]);
                    const SF vc = lae_sf(v, c);
                    SF vl;
                    vl.h = ((allowed >> j) & 1u) ? vc.h : v.h;
                    vl.l = ((allowed >> j) & 1u) ? vc.l : v.l;
                    b1[j] = add_norm(u, Kb, fb);
                    st[j] = add_norm(v, Ks_[j], fs[j]);
                    lb[j] = add_norm(vl, Kl[j], fl[j]);
                    b0[j] = add_norm(w0, Kb, fb);
                }
            }
            if (lane == 31 && w + 1 < W) mail[(i & 1) * W + w] = make_float4(lb[J - 1].h, lb[J - 1].l, 0.0f, 0.0f);
        } else {
            if (i == 0) {
                // beta at the last frame: the four final states (ha/star.py:156-163), emission included
                SF z; z.h = 0.0f; z.l = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = k0 + 32 * j;
                    if (k == L) { b0[j] = add_norm(z, Kb, fb); st[j] = add_norm(z, Ks_[j], fs[j]); b1[j] = b0[j]; }
                    if (k == L - 1) lb[j] = add_norm(z, Kl[j], fl[j]);
                }
            } else if (i >= first && i <= last) {
                // next quad's first blank and label, from the lane above (lane 31 takes lane 0 of the slot
                // above, or the mailbox of the warp above)
                float n0h[J], n0l[J], nlh[J], nll[J];
                {
                    float r0h[J], r0l[J], rlh[J], rll[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        r0h[j] = __shfl_sync(0xffffffffu, b0[j].h, (lane + 1) & 31);
                        r0l[j] = __shfl_sync(0xffffffffu, b0[j].l, (lane + 1) & 31);
                        rlh[j] = __shfl_sync(0xffffffffu, lb[j].h, (lane + 1) & 31);
                        rll[j] = __shfl_sync(0xffffffffu, lb[j].l, (lane + 1) & 31);
                    }
                    float4 in = make_float4(kVoid, 0.0f, kVoid, 0.0f);
                    if (w + 1 < W) in = mail[((i - 1) & 1) * W + w + 1];
                    cin_h = fmaxf(in.x, in.z);
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const bool edge = lane == 31;
                        const int jn = (j + 1 < J) ? j + 1 : j;
                        n0h[j] = edge ? ((j + 1 < J) ? r0h[jn] : in.x) : r0h[j];
                        n0l[j] = edge ? ((j + 1 < J) ? r0l[jn] : in.y) : r0l[j];
                        nlh[j] = edge ? ((j + 1 < J) ? rlh[jn] : in.z) : rlh[j];
                        nll[j] = edge ? ((j + 1 < J) ? rll[jn] : in.w) : rll[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    SF n0; n0.h = n0h[j]; n0.l = n0l[j];
                    SF nl; nl.h = nlh[j]; nl.l = nll[j];
                    const SF x = lae_sf(st[j], lb[j]);
                    const SF z = lae_sf(b1[j], x);
                    const SF w0 = lae_sf(b0[j], x);
                    const SF nn = lae_sf(n0, nl);
                    SF vl;
                    vl.h = ((allowed >> j) & 1u) ? nn.h : n0.h;
                    vl.l = ((allowed >> j) & 1u) ? nn.l : n0.l;
                    b0[j] = add_norm(w0, Kb, fb);
                    st[j] = add_norm(z, Ks_[j], fs[j]);
                    b1[j] = add_norm(z, Kb, fb);
                    lb[j] = add_norm(vl, Kl[j], fl[j]);
                }
            }
            if (lane == 0 && w > 0) mail[(i & 1) * W + w] = make_float4(b0[0].h, b0[0].l, lb[0].h, lb[0].l);
        }
    };
    auto step_end = [&]() {
        named_bar_sync(barid, nthr);
        if (++g == cnt) {
            named_bar_arrive(kEmptyBarrier + 4 * dir + s, nthr + 32);
            g = 0;
            if (++s == nstage) { s = 0; fpar ^= 1u; }
        }
    };
    auto smax = [&](int j) { return fmaxf(fmaxf(b0[j].h, st[j].h), fmaxf(b1[j].h, lb[j].h)); };

    // ---------------------------------------------------------------------------- phase 1 ---
    {
        int4* prow = (int4*)(tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX + JWp) + k0;
        float* hrow = tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX + w * J;
        const long long rstep = dir ? -(long long)SPX : (long long)SPX;
        for (int i = 0; i < steps1; ++i) {
            fetch(i, steps1);
            advance(i);
            // Q11.20 storage is absolute in precision, so the per-slot base only has to stay within ~2000
            // log2 units of the states that matter: reset to the slot maximum every 8th step; slots nobody
            // has reached yet take the warp's maximum
            if ((i & 7) == 0) {
                float m[J], wm = kVoid;
#pragma unroll
                for (int j = 0; j < J; ++j) { m[j] = warp_max(smax(j)); wm = fmaxf(wm, m[j]); }
                if (!(wm > kVoidTest)) wm = cin_h;          // nobody here yet: what the neighbouring warp sends
                if (wm > kVoidTest) {
#pragma unroll
                    for (int j = 0; j < J; ++j) base[j] = (m[j] > kVoidTest) ? m[j] : wm;
                }
            }
            float bsel = base[0];
#pragma unroll
            for (int j = 1; j < J; ++j) bsel = (lane == j) ? base[j] : bsel;
#pragma unroll
            for (int j = 0; j < J; ++j)
                if ((hasq >> j) & 1u)
                    prow[32 * j] = make_int4(sf_to_fix(b0[j], base[j]), sf_to_fix(st[j], base[j]),
                                             sf_to_fix(b1[j], base[j]), sf_to_fix(lb[j], base[j]));
            if (lane < J && 32 * (w * J + lane) < Q) hrow[lane] = bsel;
            prow = (int4*)((float*)prow + rstep);
            hrow += rstep;
            step_end();
        }
    }
    __threadfence();
    fence_async_all();
    cta_phase_barrier(kPhaseBarrier, (int)blockDim.x);

    // ---------------------------------------------------------------------------- phase 2 ---
    for (int i = steps1; i < Tn; ++i) {
        const int ridx = fetch(i, Tn);
        advance(i);
        const float* trow = stg + G * E + ridx * SPX;
        const int4* orow = (const int4*)(trow + JWp) + k0;       // [32 j] = other side's copy of my quad
        const float* obase = trow + w * J;                        // [j] = its slot base
        // posterior exponent of a state = [h + other base + other integer part - K] + [l + other fraction - f] - log Z
        auto expo = [&](int jj, float& xi0, float& xf0, float& xis, float& xfs, float& xi1, float& xf1,
                        float& xil, float& xfl) {
            const bool in = (hasq >> jj) & 1u;
            const int4 o = in ? orow[32 * jj] : make_int4(kFixVoid, kFixVoid, kFixVoid, kFixVoid);
            const float ob = in ? obase[jj] : 0.0f;
            float h, l;
            fix_to_parts(o.x, h, l); xi0 = ((b0[jj].h + ob) + h) - Kb;      xf0 = (b0[jj].l + l) - fb;
            fix_to_parts(o.y, h, l); xis = ((st[jj].h + ob) + h) - Ks_[jj]; xfs = (st[jj].l + l) - fs[jj];
            fix_to_parts(o.z, h, l); xi1 = ((b1[jj].h + ob) + h) - Kb;      xf1 = (b1[jj].l + l) - fb;
            fix_to_parts(o.w, h, l); xil = ((lb[jj].h + ob) + h) - Kl[jj];  xfl = (lb[jj].l + l) - fl[jj];
            if (!((hasl >> jj) & 1u)) xil = kVoid;
        };
        if (i == steps1) {
            double mx = -1.0e300;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float a, b, c, d, e, f, gg, h;
                expo(j, a, b, c, d, e, f, gg, h);
                mx = fmax(mx, fmax(fmax((double)a + (double)b, (double)c + (double)d),
                                   fmax((double)e + (double)f, (double)gg + (double)h)));
            }
            mx = warp_max_d(mx);
            if (lane == 0) redm[w] = mx;
            named_bar_sync(barid, nthr);
            for (int x = 0; x < W; ++x) mx = fmax(mx, redm[x]);
            feasible = mx > (double)kVoidTest;
            float sm = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float a, b, c, d, e, f, gg, h;
                expo(j, a, b, c, d, e, f, gg, h);
                sm += ex2f((float)((double)a + (double)b - mx)) + ex2f((float)((double)c + (double)d - mx)) +
                      ex2f((float)((double)e + (double)f - mx)) + ex2f((float)((double)gg + (double)h - mx));
            }
            sm = warp_sum(sm);
            if (lane == 0) reds[w] = sm;
            named_bar_sync(barid, nthr);
            sm = 0.0f;
            for (int x = 0; x < W; ++x) sm += reds[x];
            const double logZ2 = mx + (double)log2f(sm);
            const double fl2 = floor(logZ2);
            IZ = feasible ? (float)fl2 : 0.0f;
            fZ = feasible ? (float)(logZ2 - fl2) : 0.0f;
        }
        float* ob = (float*)stg + G * (E + SPX) + ridx * OC;
        float* ps = (float*)stg + G * (E + SPX) + G * OC;
        float bsum = 0.0f, gsum = 0.0f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = k0 + 32 * j;
            float xi0, xf0, xis, xfs, xi1, xf1, xil, xfl;
            expo(j, xi0, xf0, xis, xfs, xi1, xf1, xil, xfl);
            const float g0 = feasible ? ex2f((xi0 - IZ) + (xf0 - fZ)) : 0.0f;
            const float g1 = feasible ? ex2f((xi1 - IZ) + (xf1 - fZ)) : 0.0f;
            const float gl = feasible ? ex2f((xil - IZ) + (xfl - fZ)) : 0.0f;
            // h = gamma(star) / (P - p_y) = 2^(log2 gamma - es), es = the star's TRUE emission: without the
            // penalty and with the row shift added back (the shift cancels in gamma, not in P - p_y)
            const float h = (feasible && ((hasq >> j) & 1u))
                                ? ex2f(((xis - IZ) - ((Ks_[j] - Kp) + ct)) + ((xfs - fZ) - (fs[j] - fp))) : 0.0f;
            bsum += g0 + g1;
            gsum += h;
            if (k < Ks) ((float2*)(ob + 4))[k] = make_float2(gl, ((exclude >> j) & 1u) ? h : 0.0f);
        }
        bsum = warp_sum(bsum);
        gsum = warp_sum(gsum);
        if (lane == 0) { ps[(ridx * W + w) * 2] = bsum; ps[(ridx * W + w) * 2 + 1] = gsum; }
        if (g == cnt - 1) fence_async_smem();
        step_end();
    }
    if (dir == 0 && leader) {
        const float v = feasible ? (float)(-((double)IZ + (double)fZ + (double)csum) * kLn2) : CUDART_INF_F;
        p.loss[n] = v; p.loss_ws[n] = v;
    }
}

// ------------------------------------------------------------------------------------ grad ---
struct StarGradParams {
    const float* x; long long sx_t, sx_n;
    float* gx; long long sg_t, sg_n;
    int T, N, V, S;
    const int4* meta; const int* tgt; const int* dupnext; int Sp;
    const float* lse2; const float* occ; int E;
    const float* gout; const float* loss;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

// d loss / d x[c] = gout * ( p_c * (delta - G + H[c]) - occ_label[c] ) for c >= 1 and
// gout * (delta * p_0 - occ_blank) for the blank, with delta = 1 through the fused log-softmax
// (from_logits) and 0 at the log-prob boundary; H[c] = sum of h_k over the stars that exclude c.
// The gradient is dense over V.
template <bool VEC4>
__global__ void __launch_bounds__(kMaxRowWarps * 32) star_grad_kernel(StarGradParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nwarps;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const int L = mt.y;
    const int Ks = min(L + 1, p.S);
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x;
    const int t0 = blockIdx.x * (nw * p.rows_per_warp);
    const int V = p.V, E = p.E, nstage = p.nstage;
    uint64_t* s_bar = (uint64_t*)smem_raw;
    int* s_tgt = (int*)(smem_raw + round_up_sz((size_t)nw * nstage * 8, 128));
    int* s_nxt = s_tgt + p.Sp;
    float* s_rows = (float*)((unsigned char*)s_tgt + round_up_sz((size_t)p.Sp * 8, 128));
    if (t0 < Tn) {
        for (int k = threadIdx.x; k < Ks; k += blockDim.x) {
            s_tgt[k] = p.tgt[(size_t)n * p.Sp + k];
            s_nxt[k] = p.dupnext[(size_t)n * p.Sp + k];
        }
    }
    float* wrows = s_rows + (size_t)warp * nstage * (V + E);
    uint64_t* wbar = s_bar + warp * nstage;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&wbar[s], 1);
    mbar_init_fence();
    __syncthreads();

    int nall = 0, nreal = 0;
    if (t0 + warp < p.T) nall = min(p.rows_per_warp, (p.T - 1 - t0 - warp) / nw + 1);
    if (t0 + warp < Tn) nreal = min(nall, (Tn - 1 - t0 - warp) / nw + 1);
    const float* xb = p.x + (long long)n * p.sx_n;
    float* gb = p.gx + (long long)n * p.sg_n;
    const float g = p.gout[n];
    const float delta = p.from_logits ? 1.0f : 0.0f;
    const uint32_t occ_bytes = (uint32_t)(4 + round_up(2 * Ks, 4)) * 4u;

    auto issue = [&](int r) {
        const int stage = r % nstage;
        const int t = t0 + warp + nw * r;
        float* dst = wrows + (size_t)stage * (V + E);
        const float* src = xb + (long long)t * p.sx_t;
        const float* osrc = p.occ + ((size_t)n * p.T + t) * E;
        if (p.use_bulk) {
            if (lane == 0) {
                mbar_expect_tx(&wbar[stage], (uint32_t)V * 4u + occ_bytes);
                bulk_g2s(dst, src, (uint32_t)V * 4u, &wbar[stage]);
                bulk_g2s(dst + V, osrc, occ_bytes, &wbar[stage]);
            }
        } else {
            for (int c = lane; c < V; c += 32) dst[c] = src[c];
            for (int c = lane; c < 4 + 2 * Ks; c += 32) dst[V + c] = osrc[c];
        }
    };
    for (int r = 0; r < min(nstage - 1, nreal); ++r) issue(r);

    for (int r = 0; r < nreal; ++r) {
        const int stage = r % nstage;
        if (r + nstage - 1 < nreal) {
            if (p.use_bulk && lane == 0) bulk_wait_read<0>();
            __syncwarp();
            issue(r + nstage - 1);
        }
        if (p.use_bulk) mbar_wait(&wbar[stage], (uint32_t)(r / nstage) & 1u);
        else __syncwarp();
        float* row = wrows + (size_t)stage * (V + E);
        float* occ = row + V;
        const int t = t0 + warp + nw * r;
        const float l2 = p.lse2[(size_t)n * p.T + t];
        const float occ_blank = occ[1], G = occ[2];
        const float p0 = ex2f(fmaf(row[0], kLog2e, -l2));
        // per-class corrections, computed from the untouched logits by the first position of each
        // label chain and parked in that position's occupancy slot
        for (int k = lane; k < Ks; k += 32) {
            const int w = s_tgt[k];
            if (!(w & kNotFirst)) {
                float sl = (k < L) ? occ[4 + 2 * k] : 0.0f, sh = occ[5 + 2 * k];
                for (int j = s_nxt[k]; j >= 0; j = s_nxt[j]) { sl += (j < L) ? occ[4 + 2 * j] : 0.0f; sh += occ[5 + 2 * j]; }
                const float pc = ex2f(fmaf(row[w & kLabelMask], kLog2e, -l2));
                occ[4 + 2 * k] = g * (pc * sh - sl);
            }
        }
        __syncwarp();
        const float sc = g * (delta - G);
        if (VEC4) {
            float4* r4 = (float4*)row;
            for (int c = lane; c < (V >> 2); c += 32) {
                float4 v = r4[c];
                v.x = sc * ex2f(fmaf(v.x, kLog2e, -l2)); v.y = sc * ex2f(fmaf(v.y, kLog2e, -l2));
                v.z = sc * ex2f(fmaf(v.z, kLog2e, -l2)); v.w = sc * ex2f(fmaf(v.w, kLog2e, -l2));
                r4[c] = v;
            }
        } else {
            for (int c = lane; c < V; c += 32) row[c] = sc * ex2f(fmaf(row[c], kLog2e, -l2));
        }
        __syncwarp();
        if (lane == 0) row[0] = g * (delta * p0 - occ_blank);
        __syncwarp();
        for (int k = lane; k < Ks; k += 32) {
            const int w = s_tgt[k];
            if (!(w & kNotFirst)) row[w & kLabelMask] += occ[4 + 2 * k];
        }
        float* dstg = gb + (long long)t * p.sg_t;
        if (p.use_bulk) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(dstg, row, (uint32_t)V * 4u); bulk_commit(); }
        } else {
            __syncwarp();
            for (int c = lane; c < V; c += 32) dstg[c] = row[c];
        }
    }
    for (int r = nreal; r < nall; ++r) {
        float* dstg = gb + (long long)(t0 + warp + nw * r) * p.sg_t;
        if (VEC4) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = lane; c < (V >> 2); c += 32) ((float4*)dstg)[c] = z;
        } else {
            for (int c = lane; c < V; c += 32) dstg[c] = 0.0f;
        }
    }
    if (p.use_bulk && lane == 0) bulk_wait_all<0>();
}

}  // namespace hab

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

// emission row (floats): [0] ct (integer row shift)  [1] blank  [2] all-star P  [3] ct as an int
// [4+2k] label y_k  [5+2k] star "anything but y_k", each the probability 2^(log2 p - ct) as a plain fp32
// (emission_linear, ctc.cuh).  The occupancy row written in place (floats): [1] label + star occupancy (blanks: 1 - it)
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
        // emission exponents in float-float arithmetic from the fp32 logits, split into an integer part and a
        // fraction (emission_split, common.cuh) and stored as linear probabilities.  log2 P and the star
        // terms are (hi, lo) float pairs too.
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
        float Ka, fa;
        split2(fmaxf(lPh, kVoid), lPl, Ka, fa);
        const float Plin = emission_linear(Ka, fa);              // all-star: P_t
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            float Kb, fb;
            emission_split(row[0], l2, ct, Kb, fb);
            *(float4*)erow = make_float4(ct, emission_linear(Kb, fb), Plin, __int_as_float(__float2int_rn(ct)));
        }
        // star emission P_t - p_y (ha/star.py:4-5 logsubexp) in the linear domain: (s - e_y) * (P_t / s) with e_y the
        // very term class y contributed to the sum s, so the difference is the sum over the other classes up to
        // the rounding of s (6e-8 s) -- subtracting an independently rounded p_y would leave 2^-22 p_y behind,
        // which matters when one label holds nearly all of P_t.  Never below the smallest normal.
        const float pscale = (s > 0.0f) ? Plin / s : 0.0f;
        for (int k = lane; k < Ks; k += 32) {
            const int y = s_tgt[k];
            float Kl, fl;
            emission_split(row[y], l2, ct, Kl, fl);
            const float pl = emission_linear(Kl, fl);
            const float ey = ex2f(fmaf(row[y], kLog2e, -m2));
            const float ps = (y != 0) ? fmaxf((s - ey) * pscale, 1.1754943508222875e-38f) : Plin;
            ((float2*)(erow + 4))[k] = make_float2(pl, ps);
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
    float* tr; int SPX, JWp;   // stored row = [JWp slot bases][4*(L+1) packed words: quads]
    float* loss; float* loss_ws;
    float pen2;                // star_penalty in log2 units
    int nstage, G, W;          // ring stages; frames per stage; compute warps per sweep direction
    int dir_bytes;
};

// Same CTA shape, ring protocol and number format as ctc_trellis_kernel (one CTA per utterance, W
// compute warps + one producer warp per sweep side, meet in the middle, extended-range linear numbers).
// Both sides keep quad k in lane k%32 of slot k/32 of warp k/(32 J); beta is not a mirror image of alpha
// here (labels have no self loop, stars have a back edge), so it has its own update and its mailbox runs
// downwards.  The star state's stored word is its PRE-emission sum: h_k = gamma(star k) / (P - p_y) is then
// (my sum) x (other side's sum) x penalty / Z / 2^ct, without dividing by the star's emission.
template <int J, int W>
__global__ void __launch_bounds__(32 * (2 * W + 2)) star_trellis_kernel(StarTrellisParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = p.G, nstage = p.nstage;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool producer = warp >= 2 * W;
    const int dir = producer ? warp - 2 * W : (warp >= W);
    // the beta side takes its warps in reverse order, so each SM sub-partition hosts an alpha warp that is
    // busy early (low label pairs are reached first) next to a beta warp that is busy late
    const int w = producer ? 0 : (dir ? 2 * W - 1 - warp : warp);
    const int n = p.order[blockIdx.x];
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    // feasible iff there is a frame per label plus one between equal neighbours (labels have no self
    // loop; every stored emission is a positive number) -- decided here as in ctc_trellis_kernel
    if (mt.z || Tn == 0 || Tn < L + mt.w) {
        const float v = mt.z ? CUDART_NAN_F : CUDART_INF_F;
        if (threadIdx.x == 0) { p.loss[n] = v; p.loss_ws[n] = v; }
        return;
    }
    const int Q = L + 1, Ks = min(L + 1, p.S);
    const int E = p.E, SPX = p.SPX, JWp = p.JWp, OC = 4 + 2 * p.Sp;
    const int SF_ = trellis_stage_floats(E, SPX, OC, G, W, 64);
    // star penalty 2^pen2 = pm * 2^pe, pm in [1, 2): the mantissa multiplies the star emission, the integer
    // part goes straight into the exponent
    const float pef = floorf(p.pen2);
    const int pe = (int)pef;
    const float pm = exp2f(p.pen2 - pef);
    const float rpm = 1.0f / pm;

    unsigned char* db = smem_raw + (size_t)dir * p.dir_bytes;
    float* stages = (float*)db;
    int4* mail = (int4*)(stages + nstage * SF_);                  // [2][W]
    double* redd = (double*)(mail + 2 * W);                       // [W] Z partial sums
    int* redi = (int*)(redd + W);                                 // [W] Z partial exponent maxima
    uint64_t* full = (uint64_t*)(redi + 2 * W);
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
        trellis_producer<2, true>(stages, SF_, full, nstage, G, W, E, SPX, OC, em_base, tr_base, occ_bytes,
                                  Tn, steps1, dir, lane);
        return;
    }

    const bool leader = (w == 0 && lane == 0);
    constexpr int nthr = 32 * W;
    auto side_barrier = [&]() {
        if (dir) asm volatile("bar.sync 2, %0;" ::"n"(nthr) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"n"(nthr) : "memory");
    };
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

    // states of my quad in slot j, and the pre-emission sums they were made from
    XF b0[J], st[J], b1[J], lb[J], s0[J], ss[J], s1[J], sl[J];
#pragma unroll
    for (int j = 0; j < J; ++j)
        b0[j] = st[j] = b1[j] = lb[j] = s0[j] = ss[j] = s1[j] = sl[j] = xf_make(1.0f, kVoidE);

    float csum = 0.0f, rZ = 1.0f, rZp = 1.0f;
    double logZ2 = 0.0;
    int eZ = 0, cti = 0;
    bool feasible = true;
    float psm[J];                 // star emission x penalty mantissa of the current frame
    int s = 0; uint32_t fpar = 0;

    auto xf_norm = [](XF a) {
        const int bits = __float_as_int(a.m);
        return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), a.e + (bits >> 23) - 127);
    };
    // one step of the recursion on the emission row `er` (ha/star.py:123-145)
    auto advance = [&](const float* er, int i, bool active) {
        const float pb = er[1], pa = er[2];
        cti = ((const int*)er)[3];
        float pl[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = k0 + 32 * j;
            // quads without an entry of their own (k == L == S: the all-star, ha/star.py:47; or no quad at
            // all: a phantom of real magnitude) take the blank and the all-star emission
            const float2 wv = ((hase >> j) & 1u) ? ((const float2*)(er + 4))[k] : make_float2(pb, pa);
            pl[j] = wv.x;
            psm[j] = wv.y * pm;
        }
        if (dir == 0) {
            if (active) {
                // previous label, from the lane below (virtual state -1 holds probability 1 before the first frame)
                float cm[J]; int ce[J];
                {
                    float rm[J]; int re[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        rm[j] = __shfl_sync(0xffffffffu, lb[j].m, (lane + 31) & 31);
                        re[j] = __shfl_sync(0xffffffffu, lb[j].e, (lane + 31) & 31);
                    }
                    int4 in = make_int4(__float_as_int(1.0f), (i == 0) ? 0 : kVoidE, 0, 0);
                    if (w > 0) in = mail[((i - 1) & 1) * W + w - 1];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        cm[j] = lane ? rm[j] : (j ? rm[j ? j - 1 : 0] : __int_as_float(in.x));
                        ce[j] = lane ? re[j] : (j ? re[j ? j - 1 : 0] : in.y);
                    }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const XF c = xf_make(cm[j], ce[j]);
                    const XF u = xf_add(st[j], b1[j]);
                    const XF v = xf_add(u, b0[j]);
                    const XF w0 = xf_add(c, b0[j]);
                    const XF vc = xf_add(v, c);
                    const bool al = (allowed >> j) & 1u;
                    const XF vl = xf_make(al ? vc.m : v.m, al ? vc.e : v.e);
                    s0[j] = w0; ss[j] = v; s1[j] = u; sl[j] = vl;
                    b1[j] = xf_mul_norm(u, pb);
                    st[j] = xf_mul_norm(v, psm[j]); st[j].e += pe;
                    lb[j] = xf_mul_norm(vl, pl[j]);
                    b0[j] = xf_mul_norm(w0, pb);
                }
            }
            if (lane == 31 && w + 1 < W) mail[(i & 1) * W + w] = make_int4(__float_as_int(lb[J - 1].m), lb[J - 1].e, 0, 0);
        } else {
            if (i == 0) {
                // beta at the last frame: the four final states (ha/star.py:156-163), emission included
                const XF one = xf_make(1.0f, 0);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = k0 + 32 * j;
                    if (k == L) {
                        s0[j] = ss[j] = s1[j] = one;
                        b0[j] = xf_mul_norm(one, pb); b1[j] = b0[j];
                        st[j] = xf_mul_norm(one, psm[j]); st[j].e += pe;
                    }
                    if (k == L - 1) { sl[j] = one; lb[j] = xf_mul_norm(one, pl[j]); }
                }
            } else if (active) {
                // next quad's first blank and label, from the lane above (lane 31 takes lane 0 of the slot
                // above, or the mailbox of the warp above)
                float n0m[J], nlm[J]; int n0e[J], nle[J];
                {
                    float r0m[J], rlm[J]; int r0e[J], rle[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        r0m[j] = __shfl_sync(0xffffffffu, b0[j].m, (lane + 1) & 31);
                        r0e[j] = __shfl_sync(0xffffffffu, b0[j].e, (lane + 1) & 31);
                        rlm[j] = __shfl_sync(0xffffffffu, lb[j].m, (lane + 1) & 31);
                        rle[j] = __shfl_sync(0xffffffffu, lb[j].e, (lane + 1) & 31);
                    }
                    int4 in = make_int4(__float_as_int(1.0f), kVoidE, __float_as_int(1.0f), kVoidE);
                    if (w + 1 < W) in = mail[((i - 1) & 1) * W + w + 1];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const bool edge = lane == 31;
                        const int jn = (j + 1 < J) ? j + 1 : j;
                        n0m[j] = edge ? ((j + 1 < J) ? r0m[jn] : __int_as_float(in.x)) : r0m[j];
                        n0e[j] = edge ? ((j + 1 < J) ? r0e[jn] : in.y) : r0e[j];
                        nlm[j] = edge ? ((j + 1 < J) ? rlm[jn] : __int_as_float(in.z)) : rlm[j];
                        nle[j] = edge ? ((j + 1 < J) ? rle[jn] : in.w) : rle[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const XF n0 = xf_make(n0m[j], n0e[j]), nl = xf_make(nlm[j], nle[j]);
                    const XF x = xf_add(st[j], lb[j]);
                    const XF z = xf_add(b1[j], x);
                    const XF w0 = xf_add(b0[j], x);
                    const XF nn = xf_add(n0, nl);
                    const bool al = (allowed >> j) & 1u;
                    const XF vl = xf_make(al ? nn.m : n0.m, al ? nn.e : n0.e);
                    s0[j] = w0; ss[j] = z; s1[j] = z; sl[j] = vl;
                    b0[j] = xf_mul_norm(w0, pb);
                    st[j] = xf_mul_norm(z, psm[j]); st[j].e += pe;
                    b1[j] = xf_mul_norm(z, pb);
                    lb[j] = xf_mul_norm(vl, pl[j]);
                }
            }
            if (lane == 0 && w > 0)
                mail[(i & 1) * W + w] = make_int4(__float_as_int(b0[0].m), b0[0].e, __float_as_int(lb[0].m), lb[0].e);
        }
    };
    const int rsgn = dir ? -1 : 1;

    // ---------------------------------------------------------------------------- phase 1 ---
    {
        int4* prow = (int4*)(tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX + JWp) + k0;
        int* hrow = (int*)(tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX) + w * J + lane;
        const long long rstep = dir ? -(long long)SPX : (long long)SPX;
        const bool hstore = lane < J && 32 * (w * J + lane) < Q;
        for (int i0 = 0; i0 < steps1; i0 += G) {
            const int cnt = min(G, steps1 - i0);
            const float* er = stages + s * SF_ + (dir ? cnt - 1 : 0) * E;
            mbar_wait(&full[s], fpar);
            for (int i = i0; i < i0 + cnt; ++i, er += rsgn * E) {
                csum += er[0];
                advance(er, i, i >= first && i <= last);
                // stored relative to each slot's largest exponent of this very step; the star state's word is
                // its (normalised) pre-emission sum
                int bsel = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const XF sn = xf_norm(ss[j]);
                    const int mx = __reduce_max_sync(0xffffffffu, max(max(b0[j].e, sn.e), max(b1[j].e, lb[j].e)));
                    if ((hasq >> j) & 1u)
                        prow[32 * j] = make_int4(xf_pack(b0[j].m, mx - b0[j].e), xf_pack(sn.m, mx - sn.e),
                                                 xf_pack(b1[j].m, mx - b1[j].e), xf_pack(lb[j].m, mx - lb[j].e));
                    bsel = (lane == j) ? mx : bsel;
                }
                if (hstore) *hrow = bsel;
                prow = (int4*)((int*)prow + rstep);
                hrow += rstep;
                side_barrier();
            }
            named_bar_arrive(kEmptyBarrier + 4 * dir + s, nthr + 32);
            if (++s == nstage) { s = 0; fpar ^= 1u; }
        }
    }
    __threadfence();
    fence_async_all();
    cta_phase_barrier(kPhaseBarrier, (int)blockDim.x);

    // ---------------------------------------------------------------------------- phase 2 ---
    for (int i0 = steps1; i0 < Tn; i0 += G) {
        const int cnt = min(G, Tn - i0);
        const float* stg = stages + s * SF_;
        const int r0 = dir ? cnt - 1 : 0;
        const float* er = stg + r0 * E;
        const int* trow = (const int*)(stg + G * E + r0 * SPX);
        float* ob = (float*)stg + G * (E + SPX) + r0 * OC + 4;
        float* ps = (float*)stg + G * (E + SPX) + G * OC + (r0 * W + w) * 64 + lane;
        mbar_wait(&full[s], fpar);
        for (int i = i0; i < i0 + cnt; ++i, er += rsgn * E, trow += rsgn * SPX, ob += rsgn * OC, ps += rsgn * (W * 64)) {
            csum += er[0];
            const bool active = i >= first && i <= last;
            advance(er, i, active);
            const bool live = active || (dir == 1 && i == 0);
            const int4* orow = (const int4*)(trow + JWp) + k0;       // [32 j] = the other side's copy of my quad
            const int* obase = trow + w * J;                         // [j] = its slot base
            if (i == steps1) {
                // Z = sum over my side's states at the meeting frame of (my pre-emission sum) x (the other
                // side's value); the star states' stored words lack their emission, added here
                XF z[4 * J];
                int zm = 4 * kVoidE;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const bool in = live && ((hasq >> j) & 1u);
                    const int4 o = in ? orow[32 * j] : make_int4((int)kPackVoid, (int)kPackVoid, (int)kPackVoid, (int)kPackVoid);
                    const int ob_ = in ? obase[j] : kVoidE;
                    z[4 * j + 0] = xf_norm(xf_make(s0[j].m * xf_unpack_m(o.x), s0[j].e + ob_ - xf_unpack_below(o.x)));
                    z[4 * j + 1] = xf_mul_norm(xf_make(ss[j].m * xf_unpack_m(o.y), ss[j].e + ob_ - xf_unpack_below(o.y) + pe), psm[j]);
                    z[4 * j + 2] = xf_norm(xf_make(s1[j].m * xf_unpack_m(o.z), s1[j].e + ob_ - xf_unpack_below(o.z)));
                    z[4 * j + 3] = xf_norm(xf_make(sl[j].m * xf_unpack_m(o.w), sl[j].e + ob_ - xf_unpack_below(o.w)));
                    if (!(in && ((hasl >> j) & 1u))) z[4 * j + 3].e = 4 * kVoidE;
#pragma unroll
                    for (int c = 0; c < 4; ++c) zm = max(zm, z[4 * j + c].e);
                }
                zm = __reduce_max_sync(0xffffffffu, zm);
                if (lane == 0) redi[w] = zm;
                side_barrier();
                for (int x = 0; x < W; ++x) zm = max(zm, redi[x]);
                feasible = zm > kVoidETest;
                double sm = 0.0;
#pragma unroll
                for (int c = 0; c < 4 * J; ++c)
                    sm += (double)(z[c].m * __int_as_float((max(z[c].e - zm, -127) + 127) << 23));
#pragma unroll
                for (int o = 16; o; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                if (lane == 0) redd[w] = sm;
                side_barrier();
                sm = 0.0;
                for (int x = 0; x < W; ++x) sm += redd[x];
                const int ex = feasible ? ilogb(sm) : 0;
                rZ = feasible ? (float)(1.0 / scalbn(sm, -ex)) : 1.0f;
                rZp = rZ * pm;
                eZ = feasible ? zm + ex : (1 << 29);
                logZ2 = feasible ? (double)zm + log2(sm) : 0.0;
            }
            float bsum = 0.0f, gsum = 0.0f;
            if (live) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = k0 + 32 * j;
                    const bool in = (hasq >> j) & 1u;
                    const int4 o = in ? orow[32 * j] : make_int4((int)kPackVoid, (int)kPackVoid, (int)kPackVoid, (int)kPackVoid);
                    const int xb = (in ? obase[j] : kVoidE) - eZ;
                    float gl = xf_scale(sl[j].m * xf_unpack_m(o.w) * rZ, max(sl[j].e + xb - xf_unpack_below(o.w), -126));
                    // h = gamma(star) / (P - p_y): both sides' pre-emission sums, the penalty, and the row shift
                    // taken back out (the shift cancels in gamma, not in P - p_y)
                    const float hm = ss[j].m * xf_unpack_m(o.y) * rZp;
                    const int hx = ss[j].e + xb - xf_unpack_below(o.y) + pe;
                    const float h = xf_scale(hm, max(hx - cti, -126));
                    // gamma(star) itself = the same product times the star's (shifted) emission.  The two blank
                    // states' occupancies are not formed: a frame's occupancies sum to one, so the gradient
                    // kernel takes the blank column as 1 - (labels + stars), which this row carries in [1].
                    const float gs = hm * ((psm[j] * rpm) * __int_as_float((min(max(hx, -127), 127) + 127) << 23));
                    if (!((hasl >> j) & 1u)) gl = 0.0f;
                    bsum += in ? gl + gs : 0.0f;
                    gsum += h;
                    if (k < Ks) ((float2*)ob)[k] = make_float2(gl, ((exclude >> j) & 1u) ? h : 0.0f);
                }
            } else {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int k = k0 + 32 * j;
                    if (k < Ks) ((float2*)ob)[k] = make_float2(0.0f, 0.0f);
                }
            }
            ps[0] = bsum; ps[32] = gsum;     // per-lane partial sums; the producer warp reduces them into floats [1], [2]
            if (i == i0 + cnt - 1) fence_async_smem();
            side_barrier();
        }
        named_bar_arrive(kEmptyBarrier + 4 * dir + s, nthr + 32);
        if (++s == nstage) { s = 0; fpar ^= 1u; }
    }
    if (dir == 0 && leader) {
        const float v = feasible ? (float)(-(logZ2 + (double)csum) * kLn2) : CUDART_INF_F;
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
        const float occ_blank = 1.0f - occ[1], G = occ[2];      // occ[1] = sum of the label and star occupancies
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

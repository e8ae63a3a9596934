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

__host__ inline StarWs star_ws_layout(int T, int N, int S) {
    StarWs w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.E = round_up(4 + 2 * S, 4);                   // row shift, blank, all-star, pad, S x (label, star\label)
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

// emission row: [0] integer row shift c_t  [1] blank  [2] log2 P  [4+2k] label y_k  [5+2k] log2(P - p_{y_k})
// (log2 P when y_k == 0), all relative to c_t = rint(max(blank, log2 P)),
// for k < min(L_n + 1, S): the star in front of position L_n reads targets[n, L_n] (ha/star.py:46-47).
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
        float* erow = p.em + ((size_t)n * p.T + t) * p.E;
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            erow[0] = ct;
            erow[1] = eblank - ct;
            erow[2] = lP - ct;
        }
        for (int k = lane; k < Ks; k += 32) {
            const int y = s_tgt[k];
            const float lab = fmaf(row[y], kLog2e, -l2);
            float sub = lP;
            if (y != 0) {
                // logsubexp (ha/star.py:4-5): log2 P + log2(1 - 2^(lab - log2 P)), via expm1 so that a
                // label holding almost all of P does not cancel
                const float d = fminf(lab - lP, 0.0f);
                sub = fmaxf(lP + log2f(-expm1f(d * (float)kLn2)), kVoid);
            }
            ((float2*)(erow + 4))[k] = make_float2(lab - ct, fmaxf(sub - ct, kVoid));
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- trellis ---
__host__ __device__ inline int trellis_warp_bytes(int E, int SPX, int nstage) {
    return round_up((nstage * (E + SPX) + 2 * E) * 4 + 2 * nstage * 8, 128);
}

struct StarTrellisParams {
    int T, N, S;
    const int4* meta; const int* order; const int* tgt; int Sp;
    float* em; int E;          // emission rows in (layout above); occupancy rows out, in place:
                               // [1] blank occupancy  [2] G = sum_k h_k  [4+2k] label occupancy
                               // [5+2k] h_k (0 if y_k == 0), with h_k = gamma(star k) / (P - p_{y_k})
    float* tr; int SPX, JWp;   // stored row = [JWp slot bases][4*(L+1) floats relative to them: quads]
    float* loss; float* loss_ws;
    float pen2;                // star_penalty in log2 units
    int nstage; int warp_bytes;
};

// grid ceil(N/2), block 128: warps (2u, 2u+1) are the alpha and beta side of one utterance (see
// ctc_trellis_kernel: same meet-in-the-middle schedule, same split numbers, straight-line slots).
// Both sides keep quad k in lane k%32 of slot k/32; beta is not a mirror image of alpha here
// (labels have no self loop, stars have a back edge), so it has its own update.
template <int J>
__global__ void __launch_bounds__(128, 1) star_trellis_kernel(StarTrellisParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int usel = warp >> 1, dir = warp & 1;
    const int idx = blockIdx.x * 2 + usel;
    if (idx >= p.N) return;
    const int n = p.order[idx];
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    if (mt.z || Tn == 0) {
        const float v = mt.z ? CUDART_NAN_F : CUDART_INF_F;
        if (dir == 0 && lane == 0) { p.loss[n] = v; p.loss_ws[n] = v; }
        return;
    }
    const int Q = L + 1, Ks = min(L + 1, p.S);
    const int nstage = p.nstage, E = p.E, SPX = p.SPX, JWp = p.JWp;
    const float Kp = round_int(p.pen2), fp = p.pen2 - Kp;

    unsigned char* wb = smem_raw + (size_t)warp * p.warp_bytes;
    float* em_ring = (float*)wb;
    float* tr_ring = em_ring + nstage * E;
    float* occ_buf = tr_ring + nstage * SPX;
    uint64_t* bar_em = (uint64_t*)(occ_buf + 2 * E);
    uint64_t* bar_tr = bar_em + nstage;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) { mbar_init(&bar_em[s], 1); mbar_init(&bar_tr[s], 1); }
    mbar_init_fence();
    __syncwarp();

    // alpha: may label k be entered from label k-1?   beta: may label k be left for label k+1?
    unsigned allowed = 0, exclude = 0;   // exclude: star k excludes a class (y_k != 0)
    {
        const int* y = p.tgt + (size_t)n * p.Sp;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = 32 * j + lane;
            if (k < Ks && (y[k] & kLabelMask) != 0) exclude |= 1u << j;
            if (dir == 0) {
                if (k < L && (k == 0 || (y[k] & kLabelMask) != (y[k - 1] & kLabelMask))) allowed |= 1u << j;
            } else {
                if (k + 1 < L && (y[k + 1] & kLabelMask) != (y[k] & kLabelMask)) allowed |= 1u << j;
            }
        }
    }

    SF b0[J], st[J], b1[J], lb[J];
    float base[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { b0[j] = st[j] = b1[j] = lb[j] = sf_void(); base[j] = 0.0f; }

    float* em_base = p.em + (size_t)n * p.T * E;
    float* tr_base = p.tr + (size_t)n * p.T * SPX;
    const uint32_t em_bytes = (uint32_t)round_up(4 + 2 * Ks, 4) * 4u;
    const uint32_t tr_bytes = (uint32_t)(JWp + 4 * Q) * 4u;
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;

    auto issue_em = [&](int i) {
        const int s = i % nstage, t = dir ? Tn - 1 - i : i;
        mbar_expect_tx(&bar_em[s], em_bytes);
        bulk_g2s(em_ring + s * E, em_base + (size_t)t * E, em_bytes, &bar_em[s]);
    };
    auto issue_tr = [&](int k) {
        const int s = k % nstage, i = steps1 + k, t = dir ? Tn - 1 - i : i;
        mbar_expect_tx(&bar_tr[s], tr_bytes);
        bulk_g2s(tr_ring + s * SPX, tr_base + (size_t)t * SPX, tr_bytes, &bar_tr[s]);
    };
    auto phase_switch = [&]() {
        __threadfence();
        fence_async_all();
        named_bar_sync(1 + usel, 64);
        fence_async_all();
        if (lane == 0)
            for (int k = 0; k < min(nstage, Tn - steps1); ++k) issue_tr(k);
    };

    if (lane == 0)
        for (int i = 0; i < min(nstage, Tn); ++i) issue_em(i);

    float IZ = 0.0f, fZ = 0.0f, csum = 0.0f;
    bool feasible = true;
    int se = 0; uint32_t par = 0;
    int ts = 0; uint32_t tpar = 0;

    for (int i = 0; i < Tn; ++i) {
        if (i == steps1) phase_switch();
        const int t = dir ? Tn - 1 - i : i;
        mbar_wait(&bar_em[se], par);
        const float* er = em_ring + se * E;
        const float ct = er[0];
        csum += ct;
        const float eb = er[1], eall = er[2];
        const float Kb = round_int(eb), fb = eb - Kb;
        float Kl[J], fl[J], Ks_[J], fs[J];       // label / star emissions of my quad, split; the star's
                                                  // include the penalty paid on entering it
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = 32 * j + lane;
            float el = kVoid, es = kVoid;
            if (k < Ks) {
                const float2 e = ((const float2*)(er + 4))[k];
                el = (k < L) ? e.x : kVoid;
                es = e.y;
            } else if (k == L) {
                es = eall;                        // L == S: the last star is the all-star (ha/star.py:47)
            }
            Kl[j] = round_int(el); fl[j] = el - Kl[j];
            const float ks = round_int(es);
            Ks_[j] = fmaxf(ks + Kp, kVoid); fs[j] = (es - ks) + fp;
        }
        __syncwarp();
        if (lane == 0 && i + nstage < Tn) issue_em(i + nstage);
        if (++se == nstage) { se = 0; par ^= 1u; }

        if (dir == 0) {
            // previous label, from the lane below (virtual state -1 holds 0.0 before the first frame)
            SF c[J];
            float rh_prev = (i == 0) ? 0.0f : kVoid, rl_prev = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float rh = __shfl_sync(0xffffffffu, lb[j].h, (lane + 31) & 31);
                const float rl = __shfl_sync(0xffffffffu, lb[j].l, (lane + 31) & 31);
                c[j].h = lane ? rh : rh_prev;
                c[j].l = lane ? rl : rl_prev;
                rh_prev = rh; rl_prev = rl;
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const SF u = lae_sf(st[j], b1[j]);
                const SF v = lae_sf(u, b0[j]);
                const SF w0 = lae_sf(c[j], b0[j]);
                const SF vc = lae_sf(v, c[j]);
                SF vl;
                vl.h = ((allowed >> j) & 1u) ? vc.h : v.h;
                vl.l = ((allowed >> j) & 1u) ? vc.l : v.l;
                b1[j] = add_norm(u, Kb, fb);
                st[j] = add_norm(v, Ks_[j], fs[j]);
                lb[j] = add_norm(vl, Kl[j], fl[j]);
                b0[j] = add_norm(w0, Kb, fb);
            }
        } else if (i == 0) {
            // beta at the last frame: the four final states (ha/star.py:156-163), emission included
            SF z; z.h = 0.0f; z.l = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = 32 * j + lane;
                if (k == L) { b0[j] = add_norm(z, Kb, fb); st[j] = add_norm(z, Ks_[j], fs[j]); b1[j] = b0[j]; }
                if (k == L - 1) lb[j] = add_norm(z, Kl[j], fl[j]);
            }
        } else {
            // next quad's first blank and label, from the lane above (lane 31 takes lane 0 of the slot above)
            SF n0[J], nl[J];
            float r0h[J], r0l[J], rlh[J], rll[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                r0h[j] = __shfl_sync(0xffffffffu, b0[j].h, (lane + 1) & 31);
                r0l[j] = __shfl_sync(0xffffffffu, b0[j].l, (lane + 1) & 31);
                rlh[j] = __shfl_sync(0xffffffffu, lb[j].h, (lane + 1) & 31);
                rll[j] = __shfl_sync(0xffffffffu, lb[j].l, (lane + 1) & 31);
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const bool edge = lane == 31;
                n0[j].h = edge ? ((j + 1 < J) ? r0h[(j + 1 < J) ? j + 1 : j] : kVoid) : r0h[j];
                n0[j].l = edge ? ((j + 1 < J) ? r0l[(j + 1 < J) ? j + 1 : j] : 0.0f) : r0l[j];
                nl[j].h = edge ? ((j + 1 < J) ? rlh[(j + 1 < J) ? j + 1 : j] : kVoid) : rlh[j];
                nl[j].l = edge ? ((j + 1 < J) ? rll[(j + 1 < J) ? j + 1 : j] : 0.0f) : rll[j];
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const SF x = lae_sf(st[j], lb[j]);
                const SF z = lae_sf(b1[j], x);
                const SF w0 = lae_sf(b0[j], x);
                const SF nn = lae_sf(n0[j], nl[j]);
                SF vl;
                vl.h = ((allowed >> j) & 1u) ? nn.h : n0[j].h;
                vl.l = ((allowed >> j) & 1u) ? nn.l : n0[j].l;
                b0[j] = add_norm(w0, Kb, fb);
                st[j] = add_norm(z, Ks_[j], fs[j]);
                b1[j] = add_norm(z, Kb, fb);
                lb[j] = add_norm(vl, Kl[j], fl[j]);
            }
        }
        if (i < steps1) {
            unsigned up = 0, near = 0, live = 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float m = fmaxf(fmaxf(b0[j].h, st[j].h), fmaxf(b1[j].h, lb[j].h));
                const float d = m - base[j];
                up |= (d > kRebase) ? (1u << j) : 0u;
                near |= (d > -kRebase) ? (1u << j) : 0u;
                live |= (m > kVoidTest) ? (1u << j) : 0u;
            }
            up = __reduce_or_sync(0xffffffffu, up);
            near = __reduce_or_sync(0xffffffffu, near);
            live = __reduce_or_sync(0xffffffffu, live);
            const unsigned need = up | (live & ~near);
            if (need) {
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if ((need >> j) & 1u)
                        base[j] = warp_max(fmaxf(fmaxf(b0[j].h, st[j].h), fmaxf(b1[j].h, lb[j].h)));
            }
            float* row = tr_base + (size_t)t * SPX;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = 32 * j + lane;
                if (k < Q)
                    ((float4*)(row + JWp))[k] = make_float4((b0[j].h - base[j]) + b0[j].l, (st[j].h - base[j]) + st[j].l,
                                                            (b1[j].h - base[j]) + b1[j].l, (lb[j].h - base[j]) + lb[j].l);
                if (lane == j && 32 * j < Q) row[j] = base[j];
            }
        } else {
            const int k2 = i - steps1;
            mbar_wait(&bar_tr[ts], tpar);
            const float4* orow = (const float4*)(tr_ring + ts * SPX + JWp);
            const float* obase = tr_ring + ts * SPX;
            // posterior exponent of a state = [h + other base - K] + [l + other value - f] - log Z
            auto expo = [&](int jj, float& xi0, float& xf0, float& xis, float& xfs, float& xi1, float& xf1,
                            float& xil, float& xfl) {
                const int k = 32 * jj + lane;
                const bool in = k < Q;
                const float4 o = in ? orow[k] : make_float4(kVoid, kVoid, kVoid, kVoid);
                const float ob = in ? obase[jj] : 0.0f;
                xi0 = (b0[jj].h + ob) - Kb;      xf0 = (b0[jj].l + o.x) - fb;
                xis = (st[jj].h + ob) - Ks_[jj]; xfs = (st[jj].l + o.y) - fs[jj];
                xi1 = (b1[jj].h + ob) - Kb;      xf1 = (b1[jj].l + o.z) - fb;
                xil = (k < L) ? (lb[jj].h + ob) - Kl[jj] : kVoid;
                xfl = (lb[jj].l + o.w) - fl[jj];
            };
            if (k2 == 0) {
                double mx = -1.0e300;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float a, b, c, d, e, f, g, h;
                    expo(j, a, b, c, d, e, f, g, h);
                    mx = fmax(mx, fmax(fmax((double)a + (double)b, (double)c + (double)d),
                                       fmax((double)e + (double)f, (double)g + (double)h)));
                }
                mx = warp_max_d(mx);
                feasible = mx > (double)kVoidTest;
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float a, b, c, d, e, f, g, h;
                    expo(j, a, b, c, d, e, f, g, h);
                    s += ex2f((float)((double)a + (double)b - mx)) + ex2f((float)((double)c + (double)d - mx)) +
                         ex2f((float)((double)e + (double)f - mx)) + ex2f((float)((double)g + (double)h - mx));
                }
                s = warp_sum(s);
                const double logZ2 = mx + (double)log2f(s);
                const double fl2 = floor(logZ2);
                IZ = feasible ? (float)fl2 : 0.0f;
                fZ = feasible ? (float)(logZ2 - fl2) : 0.0f;
            }
            float* ob = occ_buf + (i & 1) * E;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            float bsum = 0.0f, gsum = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = 32 * j + lane;
                float xi0, xf0, xis, xfs, xi1, xf1, xil, xfl;
                expo(j, xi0, xf0, xis, xfs, xi1, xf1, xil, xfl);
                const float g0 = feasible ? ex2f((xi0 - IZ) + (xf0 - fZ)) : 0.0f;
                const float g1 = feasible ? ex2f((xi1 - IZ) + (xf1 - fZ)) : 0.0f;
                const float gl = feasible ? ex2f((xil - IZ) + (xfl - fZ)) : 0.0f;
                // h = gamma(star) / (P - p_y) = 2^(log2 gamma - es), es = the star's TRUE emission: without the
                // penalty and with the row shift added back (the shift cancels in gamma, not in P - p_y)
                const float h = (feasible && k < Q)
                                    ? ex2f(((xis - IZ) - ((Ks_[j] - Kp) + ct)) + ((xfs - fZ) - (fs[j] - fp))) : 0.0f;
                bsum += g0 + g1;
                gsum += h;
                if (k < Ks) ((float2*)(ob + 4))[k] = make_float2(gl, ((exclude >> j) & 1u) ? h : 0.0f);
            }
            bsum = warp_sum(bsum);
            gsum = warp_sum(gsum);
            if (lane == 0) { ob[1] = bsum; ob[2] = gsum; }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(em_base + (size_t)t * E, ob, em_bytes);
                bulk_commit();
                if (k2 + nstage < Tn - steps1) issue_tr(k2 + nstage);
            }
            if (++ts == nstage) { ts = 0; tpar ^= 1u; }
        }
    }
    if (steps1 == Tn) phase_switch();
    if (dir == 0 && lane == 0) {
        const float v = feasible ? (float)(-((double)IZ + (double)fZ + (double)csum) * kLn2) : CUDART_INF_F;
        p.loss[n] = v; p.loss_ws[n] = v;
    }
    if (lane == 0) bulk_wait_all<0>();
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
    const uint32_t occ_bytes = (uint32_t)round_up(4 + 2 * Ks, 4) * 4u;

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

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
    w.E = round_up(2 + 2 * S, 4);                   // blank, all-star, S x (label, star\label)
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

// emission row: [0] blank  [1] log2 P  [2+2k] label y_k  [3+2k] log2(P - p_{y_k}) (log2 P when y_k == 0)
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
        float* erow = p.em + ((size_t)n * p.T + t) * p.E;
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            erow[0] = fmaf(row[0], kLog2e, -l2);
            erow[1] = lP;
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
            ((float2*)(erow + 2))[k] = make_float2(lab, sub);
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- trellis ---
struct StarTrellisParams {
    int T, N, S;
    const int4* meta; const int* order; const int* tgt; int Sp;
    float* em; int E;          // emissions in; occupancy rows out (in place): [0] blank occupancy
                               // [1] G = sum_k h_k   [2+2k] label occupancy   [3+2k] h_k (0 if y_k == 0)
                               // with h_k = gamma(star k) / (P - p_{y_k})
    float* tr; int SPX, JWp;   // row = [JWp slot offsets][4*(L+1) floats: quads]
    float* loss; float* loss_ws;
    float pen2;                // star_penalty in log2 units
    int nstage; int warp_bytes;
};

// grid ceil(N/2), block 128: warps (2u, 2u+1) are the alpha and beta side of one utterance (see
// ctc_trellis_kernel).  Both sides keep quad k in lane k%32 of slot k/32; beta is not a mirror image
// of alpha here (labels have no self loop, stars have a back edge), so it has its own update.
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
    const int nslot = (Q + 31) >> 5;
    const int nstage = p.nstage, E = p.E, SPX = p.SPX, JWp = p.JWp;
    const float pen2 = p.pen2;

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

    float b0[J], st[J], b1[J], lb[J];
    int off[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { b0[j] = st[j] = b1[j] = lb[j] = kVoid; off[j] = 0; }

    float* em_base = p.em + (size_t)n * p.T * E;
    float* tr_base = p.tr + (size_t)n * p.T * SPX;
    const uint32_t em_bytes = (uint32_t)round_up(2 + 2 * Ks, 4) * 4u;
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

    int IZ = 0; float fZ = 0.0f; bool feasible = true;

    for (int i = 0; i < Tn; ++i) {
        if (i == steps1) phase_switch();
        const int s_em = i % nstage;
        const int t = dir ? Tn - 1 - i : i;
        mbar_wait(&bar_em[s_em], (uint32_t)(i / nstage) & 1u);
        const float* er = em_ring + s_em * E;
        const float eb = er[0];
        float el[J], es[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int k = 32 * j + lane;
            el[j] = kVoid; es[j] = kVoid;
            if (k < Ks) {
                const float2 e = ((const float2*)(er + 2))[k];
                el[j] = (k < L) ? e.x : kVoid;
                es[j] = e.y;
            } else if (k == L) {
                es[j] = er[1];                 // L == S: the last star is the all-star (ha/star.py:47)
            }
        }
        __syncwarp();
        if (lane == 0 && i + nstage < Tn) issue_em(i + nstage);

        if (dir == 0) {
            // previous label, from the lane below (virtual state -1 holds 0.0 before the first frame)
            float c[J];
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (j < nslot) c[j] = __shfl_sync(0xffffffffu, lb[j], (lane + 31) & 31);
            if (lane == 0) {
#pragma unroll
                for (int j = J - 1; j >= 1; --j)
                    if (j < nslot) c[j] = c[j - 1] + (float)(off[j - 1] - off[j]);
                c[0] = (i == 0) ? 0.0f : kVoid;
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const float u = lae2(st[j], b1[j]);
                    const float v = lae2(u, b0[j]);
                    const float w0 = lae2(c[j], b0[j]);
                    const float vl = ((allowed >> j) & 1u) ? lae2(v, c[j]) : v;
                    b1[j] = fmaxf(u + eb, kVoid);
                    st[j] = fmaxf(v + pen2 + es[j], kVoid);
                    lb[j] = fmaxf(vl + el[j], kVoid);
                    b0[j] = fmaxf(w0 + eb, kVoid);
                }
            }
        } else if (i == 0) {
            // beta at the last frame: the four final states (ha/star.py:156-163), emission included
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int k = 32 * j + lane;
                if (k == L) { b0[j] = eb; st[j] = fmaxf(pen2 + es[j], kVoid); b1[j] = eb; }
                if (k == L - 1) lb[j] = el[j];
            }
        } else {
            // next quad's first blank and label, from the lane above
            float n0[J], nl[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    n0[j] = __shfl_sync(0xffffffffu, b0[j], (lane + 1) & 31);
                    nl[j] = __shfl_sync(0xffffffffu, lb[j], (lane + 1) & 31);
                }
            }
            if (lane == 31) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (j < nslot) {
                        if (j + 1 < nslot && j + 1 < J) {
                            const float dd = (float)(off[j + 1 < J ? j + 1 : j] - off[j]);
                            n0[j] = n0[j + 1 < J ? j + 1 : j] + dd;
                            nl[j] = nl[j + 1 < J ? j + 1 : j] + dd;
                        } else {
                            n0[j] = kVoid; nl[j] = kVoid;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const float x = lae2(st[j], lb[j]);
                    const float z = lae2(b1[j], x);
                    const float w0 = lae2(b0[j], x);
                    const float vl = ((allowed >> j) & 1u) ? lae2(n0[j], nl[j]) : n0[j];
                    b0[j] = fmaxf(w0 + eb, kVoid);
                    st[j] = fmaxf(z + pen2 + es[j], kVoid);
                    b1[j] = fmaxf(z + eb, kVoid);
                    lb[j] = fmaxf(vl + el[j], kVoid);
                }
            }
        }
        if ((i % kRenorm) == kRenorm - 1) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const float m = warp_max(fmaxf(fmaxf(b0[j], st[j]), fmaxf(b1[j], lb[j])));
                    if (m > kVoidTest) {
                        const float k = rintf(m);
                        b0[j] -= k; st[j] -= k; b1[j] -= k; lb[j] -= k; off[j] += (int)k;
                    }
                }
            }
        }
        if (i < steps1) {
            float* row = tr_base + (size_t)t * SPX;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const int k = 32 * j + lane;
                    if (k < Q) ((float4*)(row + JWp))[k] = make_float4(b0[j], st[j], b1[j], lb[j]);
                    if (lane == j) ((int*)row)[j] = off[j];
                }
            }
        } else {
            const int k2 = i - steps1;
            const int ts = k2 % nstage;
            mbar_wait(&bar_tr[ts], (uint32_t)(k2 / nstage) & 1u);
            const float4* orow = (const float4*)(tr_ring + ts * SPX + JWp);
            const int* ooff = (const int*)(tr_ring + ts * SPX);
            float v0[J], vs[J], v1[J], vl[J];
            int ii[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                v0[j] = vs[j] = v1[j] = vl[j] = kVoid; ii[j] = 0;
                if (j < nslot) {
                    const int k = 32 * j + lane;
                    if (k < Q) {
                        const float4 o = orow[k];
                        v0[j] = b0[j] + o.x - eb;
                        vs[j] = st[j] + o.y - (pen2 + es[j]);
                        v1[j] = b1[j] + o.z - eb;
                        vl[j] = (k < L) ? lb[j] + o.w - el[j] : kVoid;
                        ii[j] = off[j] + ooff[j];
                    }
                }
            }
            if (k2 == 0) {
                double mx = -1.0e300;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (j < nslot)
                        mx = fmax(mx, (double)ii[j] + (double)fmaxf(fmaxf(v0[j], vs[j]), fmaxf(v1[j], vl[j])));
                mx = warp_max_d(mx);
                feasible = mx > (double)kVoidTest;
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (j < nslot) {
                        const float d = (float)((double)ii[j] - mx);
                        s += ex2f(v0[j] + d) + ex2f(vs[j] + d) + ex2f(v1[j] + d) + ex2f(vl[j] + d);
                    }
                s = warp_sum(s);
                const double logZ2 = mx + (double)log2f(s);
                const double fl = floor(logZ2);
                IZ = feasible ? (int)fl : 0;
                fZ = feasible ? (float)(logZ2 - fl) : 0.0f;
                if (dir == 0 && lane == 0) {
                    const float v = feasible ? (float)(-logZ2 * kLn2) : CUDART_INF_F;
                    p.loss[n] = v; p.loss_ws[n] = v;
                }
            }
            float* ob = occ_buf + (i & 1) * E;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            float bsum = 0.0f, gsum = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const int k = 32 * j + lane;
                    const float d = (float)(ii[j] - IZ) - fZ;
                    const float g0 = feasible ? ex2f(v0[j] + d) : 0.0f;
                    const float g1 = feasible ? ex2f(v1[j] + d) : 0.0f;
                    const float gl = feasible ? ex2f(vl[j] + d) : 0.0f;
                    // h = gamma(star) / (P - p_y) = 2^(log2 gamma - es)
                    const float h = (feasible && k < Q) ? ex2f(vs[j] + d - es[j]) : 0.0f;
                    bsum += g0 + g1;
                    gsum += h;
                    if (k < Ks) ((float2*)(ob + 2))[k] = make_float2(gl, ((exclude >> j) & 1u) ? h : 0.0f);
                }
            }
            bsum = warp_sum(bsum);
            gsum = warp_sum(gsum);
            if (lane == 0) { ob[0] = bsum; ob[1] = gsum; }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(em_base + (size_t)t * E, ob, em_bytes);
                bulk_commit();
                if (k2 + nstage < Tn - steps1) issue_tr(k2 + nstage);
            }
        }
    }
    if (steps1 == Tn) phase_switch();
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
    const uint32_t occ_bytes = (uint32_t)round_up(2 + 2 * Ks, 4) * 4u;

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
            for (int c = lane; c < 2 + 2 * Ks; c += 32) dst[V + c] = osrc[c];
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
        const float occ_blank = occ[0], G = occ[1];
        const float p0 = ex2f(fmaf(row[0], kLog2e, -l2));
        // per-class corrections, computed from the untouched logits by the first position of each
        // label chain and parked in that position's occupancy slot
        for (int k = lane; k < Ks; k += 32) {
            const int w = s_tgt[k];
            if (!(w & kNotFirst)) {
                float sl = (k < L) ? occ[2 + 2 * k] : 0.0f, sh = occ[3 + 2 * k];
                for (int j = s_nxt[k]; j >= 0; j = s_nxt[j]) { sl += (j < L) ? occ[2 + 2 * j] : 0.0f; sh += occ[3 + 2 * j]; }
                const float pc = ex2f(fmaf(row[w & kLabelMask], kLog2e, -l2));
                occ[2 + 2 * k] = g * (pc * sh - sl);
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
            if (!(w & kNotFirst)) row[w & kLabelMask] += occ[2 + 2 * k];
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

// rnnt_fg.cuh — joint-free ("factored") RNN-T loss for sm_100a.
//
// The reference's transducer forms its joint as a broadcast sum (ha/recognizer.py:104-114):
//     joint[n,t,u,:] = f[n,t,:] + g[n,u,:]        f = classifier(features) (N,T,V),  g = lm outputs (N,U+1,V)
// and hands the (N,T,U+1,V) tensor — 6.6 GB at BASELINE config 4, plus 6.6 GB of gradient — to the loss.
// With F[t,c] = exp(f[t,c] - max_c f[t,:]) and G[u,c] likewise, everything the loss needs is a product:
//     sum_c exp(joint[t,u,c]) = exp(mf[t] + mg[u]) * E[t,u],          E = F G^T          (T x V)(V x U1)
//     softmax[t,u,c]          = F[t,c] G[u,c] / E[t,u]
//     d loss / d f[t,c] = F[t,c] * (W G)[t,c]   - [c=0] sum_u occ_blank[t,u] - sum_u [c=y_u] occ_label[t,u]
//     d loss / d g[u,c] = G[u,c] * (W^T F)[u,c] - [c=0] sum_t occ_blank[t,u] - [c=y_u] sum_t occ_label[t,u]
// with W[t,u] = (occ_blank + occ_label)[t,u] / E[t,u].  Three small fp32 GEMMs per utterance around the
// lattice kernel of rnnt.cuh; the joint and its gradient are never materialised (algorithmic bytes drop
// from 8 V T (U+1) to 8 V (T + U + 1) per utterance).
//
//   rnnt_fg_stats_kernel    one warp per row of f and of g: row maximum, blank / label terms
//   rnnt_fg_gemm_kernel<E>  E = F G^T with exp2 applied as tiles are loaded; epilogue: blank / label arc
//                           probabilities in the lattice kernel's (mantissa, exponent) skewed layout
//   rnnt_lattice_kernel     (rnnt.cuh) alpha, beta, arc occupancies
//   rnnt_fg_gemm_kernel<DF> (W G) (.) F,   rnnt_fg_gemm_kernel<DG> (W^T F) (.) G, scaled by grad_out
//   rnnt_fg_fix_kernel      the sparse -occupancy terms (blank column, label columns; duplicate labels of a
//                           frame row are summed by the first position of their chain: deterministic)
#pragma once
#include "common.cuh"
#include "rnnt.cuh"

namespace hab {

struct RnntFgWs {
    size_t meta, tgt, nxt, loss, mf, mg, lf0, lg0, lgy, E, bl, lb, alpha, beta, occ, total;
    int Up, D;
};

__host__ inline RnntFgWs rnnt_fg_ws_layout(int N, int T, int U1) {
    RnntFgWs w;
    w.Up = round_up(U1 - 1 > 0 ? U1 - 1 : 1, 4);
    w.D = T + U1 - 1;
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Up);           // label | kNotFirst
    w.nxt = take(sizeof(int) * (size_t)N * w.Up);           // next position with the same label, or -1
    w.loss = take(sizeof(float) * (size_t)N);
    w.mf = take(sizeof(float) * (size_t)N * T);             // log2-domain row maxima
    w.mg = take(sizeof(float) * (size_t)N * U1);
    w.lf0 = take(sizeof(float) * (size_t)N * T);            // f[t,0] log2e - mf[t]
    w.lg0 = take(sizeof(float) * (size_t)N * U1);
    w.lgy = take(sizeof(float) * (size_t)N * U1);           // g[u,y_u] log2e - mg[u]
    w.E = take(sizeof(float) * (size_t)N * T * U1);
    w.bl = take(sizeof(float2) * (size_t)N * w.D * U1);
    w.lb = take(sizeof(float2) * (size_t)N * w.D * U1);
    w.alpha = take(sizeof(double) * (size_t)N * w.D * U1);
    w.beta = take(sizeof(double) * (size_t)N * w.D * U1);
    w.occ = take(sizeof(float2) * (size_t)N * w.D * U1);
    w.total = o;
    return w;
}

struct RnntFgParams {
    const float* f; const float* g;        // (N,T,V), (N,U1,V) contiguous
    float* gf; float* gg;                  // gradients, same shapes (backward only)
    int N, T, U1, V, Up, D;
    const int4* meta; int* tgt; int* nxt;
    float* mf; float* mg; float* lf0; float* lg0; float* lgy; float* E;
    float2* bl; float2* lb; const float2* occ;
    const float* gout; const float* loss;
};

// grid N, block 128, dynamic smem Up ints: duplicate-label chains of the targets (for the deterministic
// label fix-up): nxt[k] = next position holding the same label, tgt[k] |= kNotFirst if an earlier one does
__global__ void __launch_bounds__(128) rnnt_fg_chain_kernel(RnntFgParams p) {
    extern __shared__ int s_lab[];
    const int n = blockIdx.x;
    const int4 mt = p.meta[n];
    const int U = mt.z ? 0 : mt.y;
    int* y = p.tgt + (size_t)n * p.Up;
    int* nx = p.nxt + (size_t)n * p.Up;
    for (int k = threadIdx.x; k < U; k += blockDim.x) s_lab[k] = y[k] & kLabelMask;
    __syncthreads();
    for (int k = threadIdx.x; k < U; k += blockDim.x) {
        const int yk = s_lab[k];
        int nxt = -1, notfirst = 0;
        for (int j = k + 1; j < U; ++j) if (s_lab[j] == yk) { nxt = j; break; }
        for (int j = 0; j < k; ++j) if (s_lab[j] == yk) { notfirst = 1; break; }
        nx[k] = nxt;
        y[k] = yk | (notfirst ? kNotFirst : 0);
    }
}

// grid (ceil((T + U1) / 8), N), block 256: one warp per row of f (rows [0,T)) or g (rows [T, T+U1))
__global__ void __launch_bounds__(256) rnnt_fg_stats_kernel(RnntFgParams p) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= p.T + p.U1) return;
    const bool isg = r >= p.T;
    const int row = isg ? r - p.T : r;
    const float* x = (isg ? p.g + ((size_t)n * p.U1 + row) * p.V : p.f + ((size_t)n * p.T + row) * p.V);
    float mx = -CUDART_INF_F;
    for (int c = lane; c < p.V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    if (lane == 0) {
        const float m2 = mx * kLog2e;
        const float l0 = fmaf(x[0], kLog2e, -m2);
        if (isg) {
            const int4 mt = p.meta[n];
            p.mg[(size_t)n * p.U1 + row] = m2;
            p.lg0[(size_t)n * p.U1 + row] = l0;
            const int y = (row < mt.y) ? (p.tgt[(size_t)n * p.Up + row] & kLabelMask) : 0;
            p.lgy[(size_t)n * p.U1 + row] = fmaf(x[y], kLog2e, -m2);
        } else {
            p.mf[(size_t)n * p.T + row] = m2;
            p.lf0[(size_t)n * p.T + row] = l0;
        }
    }
}

// ------------------------------------------------------------------------------------ GEMMs ---
// C[m, n] = sum_k A(m, k) B(k, n), 64 x 64 tile per CTA, 16-wide k chunks, 256 threads x (4 x 4) outputs,
// fp32 FFMA (the operands are exponentials of fp32 logits and the result feeds a 1e-5 gradient: bf16/tf32
// tensor-core inputs would not hold that).  MODE selects operands and epilogue:
//   kE : A = F (T x V),  B = G^T (V x U1)          -> E, blank / label arc probabilities
//   kDF: A = W (T x U1), B = G   (U1 x V)          -> gf = gout F (.) (W G)
//   kDG: A = W^T (U1 x T), B = F (T x V)           -> gg = gout G (.) (W^T F)
enum { kE = 0, kDF = 1, kDG = 2 };
constexpr int kGM = 64, kGN = 64, kGK = 16;

template <int MODE>
__global__ void __launch_bounds__(256, 3) rnnt_fg_gemm_kernel(RnntFgParams p) {
    __shared__ float As[2][kGK][kGM + 4];
    __shared__ float Bs[2][kGK][kGN + 4];
    const int n = blockIdx.z;
    const int4 mt = p.meta[n];
    const float lossn = (MODE == kE) ? 0.0f : p.loss[n];
    const int Tn = (mt.z || (MODE != kE && !(lossn < CUDART_INF_F))) ? 0 : mt.x, Un = mt.y;
    const int T = p.T, U1 = p.U1, V = p.V;
    const int M = (MODE == kDG) ? U1 : T;
    const int m0 = blockIdx.x * kGM, n0 = blockIdx.y * kGN;
    const float* __restrict__ f = p.f + (size_t)n * T * V;
    const float* __restrict__ g = p.g + (size_t)n * U1 * V;
    const float* __restrict__ mf = p.mf + (size_t)n * T;
    const float* __restrict__ mg = p.mg + (size_t)n * U1;
    const float* __restrict__ Em = p.E + (size_t)n * T * U1;
    const float2* __restrict__ occ = p.occ + (size_t)n * p.D * U1;
    const int tid = threadIdx.x;
    // valid extents of the operands (everything outside contributes 0)
    const int Mv = (MODE == kDG) ? Un + 1 : Tn;              // rows of C that matter
    const int Kv = (MODE == kE) ? V : (MODE == kDF ? Un + 1 : Tn);
    if (m0 >= Mv && MODE == kE) return;

    auto Fx = [&](int t, int c) { return ex2f(fmaf(f[(size_t)t * V + c], kLog2e, -mf[t])); };
    auto Gx = [&](int u, int c) { return ex2f(fmaf(g[(size_t)u * V + c], kLog2e, -mg[u])); };

    // Operand elements are fetched RAW into registers one k-chunk ahead (the loads are in flight while the
    // current chunk is multiplied), then converted -- exp2 of the shifted logit, or node occupancy / E --
    // as they are committed to the other shared-memory buffer.  An element outside the valid extent is
    // fetched as (-inf, 0) / (0, 1), which converts to exactly 0.
    float ax[4], ay[4], bx[4], by[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + 256 * i;
            if (MODE == kE) {            // A(m=t, k=c): c contiguous;  B(k=c, n=u): c contiguous
                const int kk = e & 15, mm = e >> 4;
                const int t = m0 + mm, c = k0 + kk, u = n0 + mm;
                const bool va = t < Tn && c < V, vb = u <= Un && c < V;
                ax[i] = va ? f[(size_t)t * V + c] : -CUDART_INF_F; ay[i] = va ? mf[t] : 0.0f;
                bx[i] = vb ? g[(size_t)u * V + c] : -CUDART_INF_F; by[i] = vb ? mg[u] : 0.0f;
            } else if (MODE == kDF) {    // A(m=t, k=u): u contiguous;  B(k=u, n=c): c contiguous
                const int kk = e & 15, mm = e >> 4;
                const int t = m0 + mm, u = k0 + kk;
                const bool va = t < Tn && u <= Un;
                const float2 o = va ? occ[(size_t)(t + u) * U1 + u] : make_float2(0.0f, 0.0f);
                ax[i] = o.x + o.y; ay[i] = va ? Em[(size_t)t * U1 + u] : 1.0f;
                const int nn = e & 63, k2 = e >> 6;
                const int u2 = k0 + k2, c = n0 + nn;
                const bool vb = u2 <= Un && c < V;
                bx[i] = vb ? g[(size_t)u2 * V + c] : -CUDART_INF_F; by[i] = vb ? mg[u2] : 0.0f;
            } else {                     // A(m=u, k=t): u contiguous;  B(k=t, n=c): c contiguous
                const int mm = e & 63, kk = e >> 6;
                const int u = m0 + mm, t = k0 + kk, c = n0 + mm;
                const bool va = t < Tn && u <= Un, vb = t < Tn && c < V;
                const float2 o = va ? occ[(size_t)(t + u) * U1 + u] : make_float2(0.0f, 0.0f);
                ax[i] = o.x + o.y; ay[i] = va ? Em[(size_t)t * U1 + u] : 1.0f;
                bx[i] = vb ? f[(size_t)t * V + c] : -CUDART_INF_F; by[i] = vb ? mf[t] : 0.0f;
            }
        }
    };
    auto commit = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + 256 * i;
            const float av = (MODE == kE) ? ex2f(fmaf(ax[i], kLog2e, -ay[i])) : ((ay[i] > 0.0f) ? ax[i] / ay[i] : 0.0f);
            const float bv = ex2f(fmaf(bx[i], kLog2e, -by[i]));
            if (MODE == kE) {
                As[buf][e & 15][e >> 4] = av; Bs[buf][e & 15][e >> 4] = bv;
            } else if (MODE == kDF) {
                As[buf][e & 15][e >> 4] = av; Bs[buf][e >> 6][e & 63] = bv;
            } else {
                As[buf][e >> 6][e & 63] = av; Bs[buf][e >> 6][e & 63] = bv;
            }
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;      // my 4 x 4 outputs inside the tile

    int buf = 0;
    if (Kv > 0) { fetch(0); commit(0); }
    __syncthreads();
    for (int k0 = 0; k0 < Kv; k0 += kGK) {
        const bool more = k0 + kGK < Kv;
        if (more) fetch(k0 + kGK);
        // two-level accumulation: the 16 products of a chunk are summed on their own and added to the running
        // total once, so the total is rounded K/16 times instead of K times (the blank column of gg is a
        // difference of two sums over T frames)
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < kGK; ++kk) {
            const float4 a = *(const float4*)&As[buf][kk][tm];
            const float4 b = *(const float4*)&Bs[buf][kk][tn];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
        if (more) commit(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    if (MODE == kE) {
        const int* y = p.tgt + (size_t)n * p.Up;
        const float* lf0 = p.lf0 + (size_t)n * T;
        const float* lg0 = p.lg0 + (size_t)n * U1;
        const float* lgy = p.lgy + (size_t)n * U1;
        float2* bl = p.bl + (size_t)n * p.D * U1;
        float2* lb = p.lb + (size_t)n * p.D * U1;
        float* Eo = p.E + (size_t)n * T * U1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = m0 + tm + i;
            if (t >= Tn) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = n0 + tn + j;
                if (u > Un) continue;
                const float E = acc[i][j];
                Eo[(size_t)t * U1 + u] = E;
                const float lE = log2f(E);
                const size_t sk = (size_t)(t + u) * U1 + u;
                bl[sk] = log2_to_parts(lf0[t] + lg0[u] - lE);
                float ll = kVoid;
                if (u < Un) ll = fmaf(f[(size_t)t * V + (y[u] & kLabelMask)], kLog2e, -mf[t]) + lgy[u] - lE;
                lb[sk] = log2_to_parts(ll);
            }
        }
    } else {
        const float go = p.gout[n];
        float* out = (MODE == kDF) ? p.gf + (size_t)n * T * V : p.gg + (size_t)n * U1 * V;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + tm + i;
            if (m >= M) continue;
            const bool live = (MODE == kDF) ? (m < Tn) : (m <= Un && Tn > 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = n0 + tn + j;
                if (c >= V) continue;
                float v = 0.0f;
                if (live) v = go * acc[i][j] * ((MODE == kDF) ? Fx(m, c) : Gx(m, c));
                out[(size_t)m * V + c] = v;
            }
        }
    }
}

// grid (ceil((T + U1) / 8), N), block 256: the sparse -occupancy terms, one warp per row of gf (rows [0,T))
// or of gg (rows [T, T+U1)); runs after the two gradient GEMMs
__global__ void __launch_bounds__(256) rnnt_fg_fix_kernel(RnntFgParams p) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= p.T + p.U1) return;
    const int4 mt = p.meta[n];
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x, Un = mt.y;
    const int U1 = p.U1, V = p.V;
    const float2* occ = p.occ + (size_t)n * p.D * U1;
    const int* y = p.tgt + (size_t)n * p.Up;
    const int* nx = p.nxt + (size_t)n * p.Up;
    const float go = p.gout[n];
    if (r < p.T) {
        const int t = r;
        if (t >= Tn) return;
        float* row = p.gf + ((size_t)n * p.T + t) * V;
        float s0 = 0.0f;                    // everything that lands on class 0: blank arcs, and label arcs of label 0
        for (int u = lane; u <= Un; u += 32) s0 += occ[(size_t)(t + u) * U1 + u].x;
        // label columns: the first position of every chain of equal labels sums its chain (one writer per class)
        for (int u = lane; u < Un; u += 32) {
            const int w = y[u];
            if (!(w & kNotFirst)) {
                float s = occ[(size_t)(t + u) * U1 + u].y;
                for (int j = nx[u]; j >= 0; j = nx[j]) s += occ[(size_t)(t + j) * U1 + j].y;
                if ((w & kLabelMask) != 0) row[w & kLabelMask] -= go * s;
                else s0 += s;
            }
        }
        s0 = warp_sum(s0);
        if (lane == 0) row[0] -= go * s0;
    } else {
        const int u = r - p.T;
        if (u > Un || Tn == 0) return;
        float* row = p.gg + ((size_t)n * U1 + u) * V;
        float cb = 0.0f, cl = 0.0f;
        for (int t = lane; t < Tn; t += 32) {
            const float2 o = occ[(size_t)(t + u) * U1 + u];
            cb += o.x; cl += o.y;
        }
        cb = warp_sum(cb); cl = warp_sum(cl);
        if (lane == 0) {
            row[0] -= go * cb;
            if (u < Un) row[y[u] & kLabelMask] -= go * cl;       // lands on class 0 as well if y_u == 0
        }
    }
}

}  // namespace hab

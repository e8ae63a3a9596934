// rnnt.cuh — RNN-T lattice loss + joint-logit gradient for sm_100a.  Replaces
// ha/transducer.py:175-205 (transducer_forward_score; its column scans ha/scan.py:88-126 become an
// anti-diagonal sweep) and the autograd backward.
//
//   rnnt_prep_kernel     lengths/targets -> int32 metadata
//   rnnt_rows_kernel     one warp per lattice node (n,t,u): row log-sum-exp + the blank and label
//                        log-probs, written in a diagonal-major ("skewed") layout
//   rnnt_lattice_kernel  one CTA per utterance, one thread per u: alpha swept along anti-diagonals,
//                        then beta swept back with the arc occupancies produced on the fly; float64
//                        accumulators with fp32 MUFU for the log1p(exp) correction (the lattice is
//                        ~V times smaller than the joint, so this costs nothing)
//   rnnt_grad_kernel     one warp per node: softmax * occ_node - occ_blank - occ_label, in place in
//                        shared memory between a bulk load and a bulk store
//   rnnt_zero_kernel     zero gradient rows of padded nodes (t >= T_n or u > U_n)
#pragma once
#include "common.cuh"

namespace hab {

struct RnntWs {
    size_t meta, tgt, loss, lse2, bl, lb, alpha, beta, occ, total;
    int Up, D;   // padded target stride; diagonals per utterance
};

__host__ inline RnntWs rnnt_ws_layout(int N, int T, int U1) {
    RnntWs w;
    w.Up = round_up(U1 - 1 > 0 ? U1 - 1 : 1, 4);
    w.D = T + U1 - 1;
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Up);
    w.loss = take(sizeof(float) * (size_t)N);
    w.lse2 = take(sizeof(float) * (size_t)N * T * U1);          // row-major (t,u)
    w.bl = take(sizeof(float) * (size_t)N * w.D * U1);          // skewed (t+u, u)
    w.lb = take(sizeof(float) * (size_t)N * w.D * U1);
    w.alpha = take(sizeof(double) * (size_t)N * w.D * U1);
    w.beta = take(sizeof(double) * (size_t)N * w.D * U1);
    w.occ = take(sizeof(float2) * (size_t)N * w.D * U1);        // (occ_blank, occ_label), skewed
    w.total = o;
    return w;
}

struct RnntPrepParams {
    const void* targets; long long tgt_stride; int tgt64;
    const void* in_len; const void* tgt_len; int len64;
    int N, T, U, V, Up;
    int4* meta; int* tgt;
};

// grid N, block 128
__global__ void __launch_bounds__(128) rnnt_prep_kernel(RnntPrepParams p) {
    __shared__ int s_bad;
    const int n = blockIdx.x;
    const long long Tn = load_idx(p.in_len, n, p.len64), Un = load_idx(p.tgt_len, n, p.len64);
    if (threadIdx.x == 0) s_bad = (Tn <= 0 || Tn > p.T || Un < 0 || Un > p.U) ? 1 : 0;
    __syncthreads();
    const int U = s_bad ? 0 : (int)Un;
    for (int k = threadIdx.x; k < p.U; k += blockDim.x) {
        long long y = load_idx(p.targets, (long long)n * p.tgt_stride + k, p.tgt64);
        if (y < 0 || y >= p.V) { if (k < U) s_bad = 1; y = 0; }
        p.tgt[(size_t)n * p.Up + k] = (int)y;
    }
    __syncthreads();
    if (threadIdx.x == 0) p.meta[n] = make_int4(s_bad ? 0 : (int)Tn, U, s_bad, 0);
}

// ------------------------------------------------------------------------------------ rows ---
struct RnntRowsParams {
    const float* x;            // (N,T,U1,V) contiguous
    int N, T, U1, V;
    const int4* meta; const int* tgt; int Up;
    float* lse2; float* bl; float* lb; int D;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

__host__ __device__ inline size_t rnnt_rows_smem_bytes(int V, int nstage, int nwarps) {
    return round_up_sz((size_t)nwarps * nstage * 8, 128) + (size_t)nwarps * nstage * V * 4;
}

// grid (ceil(T*U1 / (nwarps*rows_per_warp)), N), block 32*nwarps.  Only real nodes are visited:
// node v of utterance n is (t,u) = (v / (U_n+1), v % (U_n+1)).
template <bool VEC4>
__global__ void __launch_bounds__(256) rnnt_rows_kernel(RnntRowsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nwarps, nstage = p.nstage, V = p.V;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const int Tn = mt.z ? 0 : mt.x, W = mt.y + 1;
    const int nvalid = Tn * W;
    const int v0 = blockIdx.x * (nw * p.rows_per_warp);
    if (v0 >= nvalid) return;
    uint64_t* wbar = (uint64_t*)smem_raw + warp * nstage;
    float* wrows = (float*)(smem_raw + round_up_sz((size_t)nw * nstage * 8, 128)) + (size_t)warp * nstage * V;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&wbar[s], 1);
    mbar_init_fence();
    __syncwarp();

    int nrows = 0;
    if (v0 + warp < nvalid) nrows = min(p.rows_per_warp, (nvalid - 1 - v0 - warp) / nw + 1);
    const float* xb = p.x + (size_t)n * p.T * p.U1 * V;
    const int* y = p.tgt + (size_t)n * p.Up;

    auto issue = [&](int r) {
        const int stage = r % nstage;
        const int v = v0 + warp + nw * r;
        const int t = v / W, u = v - t * W;
        const float* src = xb + ((size_t)t * p.U1 + u) * V;
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
        const int v = v0 + warp + nw * r;
        const int t = v / W, u = v - t * W;
        float l2 = 0.0f;
        if (p.from_logits) {
            float mx = -CUDART_INF_F;
            if (VEC4) {
                const float4* r4 = (const float4*)row;
                for (int c = lane; c < (V >> 2); c += 32) {
                    float4 q = r4[c];
                    mx = fmaxf(mx, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
                }
            } else {
                for (int c = lane; c < V; c += 32) mx = fmaxf(mx, row[c]);
            }
            mx = warp_max(mx);
            const float m2 = mx * kLog2e;
            float s = 0.0f;
            if (VEC4) {
                const float4* r4 = (const float4*)row;
                for (int c = lane; c < (V >> 2); c += 32) {
                    float4 q = r4[c];
                    s += ex2f(fmaf(q.x, kLog2e, -m2)) + ex2f(fmaf(q.y, kLog2e, -m2)) +
                         ex2f(fmaf(q.z, kLog2e, -m2)) + ex2f(fmaf(q.w, kLog2e, -m2));
                }
            } else {
                for (int c = lane; c < V; c += 32) s += ex2f(fmaf(row[c], kLog2e, -m2));
            }
            s = warp_sum(s);
            l2 = m2 + log2f(s);
        }
        if (lane == 0) {
            const size_t sk = ((size_t)n * p.D + (t + u)) * p.U1 + u;
            p.lse2[((size_t)n * p.T + t) * p.U1 + u] = l2;
            p.bl[sk] = fmaf(row[0], kLog2e, -l2);
            p.lb[sk] = (u < W - 1) ? fmaf(row[y[u]], kLog2e, -l2) : kVoid;
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- lattice ---
struct RnntLatticeParams {
    int N, T, U1, D;
    const int4* meta;
    const float* bl; const float* lb;
    double* alpha; double* beta; float2* occ;
    float* loss; float* loss_ws;
};

// log2(2^a + 2^b) with float64 accumulators; only the [0,1] correction term is fp32
__device__ __forceinline__ double lae2_d(double a, double b) {
    const double m = fmax(a, b);
    const float d = (float)(fmin(a, b) - m);
    return m + (double)lg2f(1.0f + ex2f(d));
}

constexpr double kVoidD = -1.0e30;

// grid N, block sides * round_up(U1, 32).  Thread u of a side owns column u; on anti-diagonal d it is
// at t = d - u.  With sides == 2 the alpha sweep (threads [0, half)) and the beta sweep (threads
// [half, 2 half)) run concurrently, each behind its own named barrier; with sides == 1 (U+1 > 512)
// the same threads run them one after the other.  Arc occupancies are then one parallel pass.
//   alpha[t,u] = (alpha[t-1,u] + blank[t-1,u]) (+) (alpha[t,u-1] + label[t,u-1])   ha/transducer.py:197-202
__global__ void __launch_bounds__(1024) rnnt_lattice_kernel(RnntLatticeParams p) {
    __shared__ double s_edge[2][2][32];     // [side][diagonal parity][warp]
    __shared__ double s_logz;
    const int n = blockIdx.x;
    const int U1 = p.U1;
    const int half = round_up(U1, 32);
    const int sides = blockDim.x / half;
    const int side = threadIdx.x / half;
    const int u = threadIdx.x - side * half, lane = u & 31, warp = u >> 5, nwarp = half >> 5;
    const int4 mt = p.meta[n];
    const int Tn = mt.x, Un = mt.y;
    if (mt.z) { if (threadIdx.x == 0) { p.loss[n] = CUDART_NAN_F; p.loss_ws[n] = CUDART_NAN_F; } return; }
    const float* bl = p.bl + (size_t)n * p.D * U1;
    const float* lb = p.lb + (size_t)n * p.D * U1;
    double* al = p.alpha + (size_t)n * p.D * U1;
    double* be = p.beta + (size_t)n * p.D * U1;
    float2* occ = p.occ + (size_t)n * p.D * U1;
    const int nd = Tn + Un;                 // diagonals 0 .. nd-1
    const bool col = u <= Un;

    if (side == 0) {
        // ---- alpha, forward over diagonals; emissions of the next diagonal are fetched a step ahead
        double a = kVoidD;                  // alpha of my node on the previous diagonal
        float nbl = 0.0f, nlb = 0.0f;       // blank[t-1,u], label[t,u-1] for the coming diagonal
        for (int d = 0; d < nd; ++d) {
            const int t = d - u;
            const float cbl = nbl, clb = nlb;
            if (d + 1 < nd && col) {
                nbl = bl[(size_t)d * U1 + u];
                nlb = (u >= 1) ? lb[(size_t)d * U1 + u - 1] : 0.0f;
            }
            double left = __shfl_up_sync(0xffffffffu, a, 1);       // alpha[t, u-1]
            if (lane == 0) left = (warp > 0) ? s_edge[0][(d + 1) & 1][warp - 1] : kVoidD;
            double cur = kVoidD;
            if (col && t >= 0 && t < Tn) {
                if (d == 0) cur = 0.0;      // alpha[0,0]
                else {
                    const double up = (t >= 1) ? a + (double)cbl : kVoidD;
                    const double lf = (u >= 1) ? left + (double)clb : kVoidD;
                    cur = lae2_d(lf, up);
                }
                al[(size_t)d * U1 + u] = cur;
            }
            a = cur;
            if (lane == 31) s_edge[0][d & 1][warp] = a;
            named_bar_sync(1, half);
        }
        // log Z = alpha[T-1,U] + blank[T-1,U]                                  ha/transducer.py:204-205
        if (u == Un) s_logz = a + (double)bl[(size_t)(nd - 1) * U1 + Un];
    }
    if (side == sides - 1) {
        // ---- beta, backward over diagonals
        double b = kVoidD;                  // beta[t+1,u]: my node on the next diagonal
        float nbl = 0.0f, nlb = 0.0f;
        if (col && nd >= 1) { nbl = bl[(size_t)(nd - 1) * U1 + u]; nlb = lb[(size_t)(nd - 1) * U1 + u]; }
        for (int d = nd - 1; d >= 0; --d) {
            const int t = d - u;
            const float cbl = nbl, clb = nlb;
            if (d >= 1 && col) { nbl = bl[(size_t)(d - 1) * U1 + u]; nlb = lb[(size_t)(d - 1) * U1 + u]; }
            double right = __shfl_down_sync(0xffffffffu, b, 1);    // beta[t, u+1]
            if (lane == 31) right = (warp + 1 < nwarp) ? s_edge[1][(d + 1) & 1][warp + 1] : kVoidD;
            double cur = kVoidD;
            if (col && t >= 0 && t < Tn) {
                double tb, tl;
                if (t == Tn - 1) tb = (u == Un) ? (double)cbl : kVoidD;   // only the terminal blank leaves the last frame
                else tb = b + (double)cbl;
                tl = (u < Un) ? right + (double)clb : kVoidD;
                cur = lae2_d(tb, tl);
                be[(size_t)d * U1 + u] = cur;
            }
            b = cur;
            if (lane == 0) s_edge[1][d & 1][warp] = b;
            named_bar_sync(2, half);
        }
    }
    __syncthreads();
    const double logz = s_logz;
    const bool feasible = logz > -1.0e29;
    if (threadIdx.x == 0) {
        const float v = feasible ? (float)(-logz * kLn2) : CUDART_INF_F;
        p.loss[n] = v; p.loss_ws[n] = v;
    }
    // ---- arc occupancies: occ_blank = alpha + blank + beta[t+1,u] - logZ, occ_label = alpha + label + beta[t,u+1] - logZ
    for (int k = threadIdx.x; k < nd * U1; k += blockDim.x) {
        const int d = k / U1, uu = k - d * U1, t = d - uu;
        if (uu > Un || t < 0 || t >= Tn) continue;
        float2 o = make_float2(0.0f, 0.0f);
        if (feasible) {
            const double base = al[k] - logz;
            double tb, tl;
            if (t == Tn - 1) tb = (uu == Un) ? (double)bl[k] : kVoidD;
            else tb = be[(size_t)(d + 1) * U1 + uu] + (double)bl[k];
            tl = (uu < Un) ? be[(size_t)(d + 1) * U1 + uu + 1] + (double)lb[k] : kVoidD;
            o.x = ex2f((float)fmax(base + tb, -200.0));
            o.y = ex2f((float)fmax(base + tl, -200.0));
        }
        occ[k] = o;
    }
}

// ------------------------------------------------------------------------------------ grad ---
struct RnntGradParams {
    const float* x; float* gx;
    int N, T, U1, V;
    const int4* meta; const int* tgt; int Up;
    const float* lse2; const float2* occ; int D;
    const float* gout; const float* loss;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

// same node enumeration as rnnt_rows_kernel
template <bool VEC4>
__global__ void __launch_bounds__(256) rnnt_grad_kernel(RnntGradParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = p.nwarps, nstage = p.nstage, V = p.V;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x, W = mt.y + 1;
    const int nvalid = Tn * W;
    const int v0 = blockIdx.x * (nw * p.rows_per_warp);
    if (v0 >= nvalid) return;
    uint64_t* wbar = (uint64_t*)smem_raw + warp * nstage;
    float* wrows = (float*)(smem_raw + round_up_sz((size_t)nw * nstage * 8, 128)) + (size_t)warp * nstage * V;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&wbar[s], 1);
    mbar_init_fence();
    __syncwarp();

    int nrows = 0;
    if (v0 + warp < nvalid) nrows = min(p.rows_per_warp, (nvalid - 1 - v0 - warp) / nw + 1);
    const size_t ubase = (size_t)n * p.T * p.U1;
    const int* y = p.tgt + (size_t)n * p.Up;
    const float g = p.gout[n];

    auto node = [&](int r, int& t, int& u) { const int v = v0 + warp + nw * r; t = v / W; u = v - t * W; };
    auto issue = [&](int r) {
        if (!p.from_logits) return;
        const int stage = r % nstage;
        int t, u; node(r, t, u);
        const float* src = p.x + (ubase + (size_t)t * p.U1 + u) * V;
        if (p.use_bulk) {
            if (lane == 0) {
                mbar_expect_tx(&wbar[stage], (uint32_t)V * 4u);
                bulk_g2s(wrows + (size_t)stage * V, src, (uint32_t)V * 4u, &wbar[stage]);
            }
        } else {
            for (int c = lane; c < V; c += 32) wrows[(size_t)stage * V + c] = src[c];
        }
    };
    for (int r = 0; r < min(nstage - 1, nrows); ++r) issue(r);

    for (int r = 0; r < nrows; ++r) {
        const int stage = r % nstage;
        if (r + nstage - 1 < nrows) {
            if (p.use_bulk && lane == 0) bulk_wait_read<0>();
            __syncwarp();
            issue(r + nstage - 1);
        }
        if (p.use_bulk && p.from_logits) mbar_wait(&wbar[stage], (uint32_t)(r / nstage) & 1u);
        else __syncwarp();
        float* row = wrows + (size_t)stage * V;
        int t, u; node(r, t, u);
        const float2 o = p.occ[((size_t)n * p.D + (t + u)) * p.U1 + u];
        if (p.from_logits) {
            const float l2 = p.lse2[ubase + (size_t)t * p.U1 + u];
            const float sc = g * (o.x + o.y);                    // node occupancy
            if (VEC4) {
                float4* r4 = (float4*)row;
                for (int c = lane; c < (V >> 2); c += 32) {
                    float4 q = r4[c];
                    q.x = sc * ex2f(fmaf(q.x, kLog2e, -l2)); q.y = sc * ex2f(fmaf(q.y, kLog2e, -l2));
                    q.z = sc * ex2f(fmaf(q.z, kLog2e, -l2)); q.w = sc * ex2f(fmaf(q.w, kLog2e, -l2));
                    r4[c] = q;
                }
            } else {
                for (int c = lane; c < V; c += 32) row[c] = sc * ex2f(fmaf(row[c], kLog2e, -l2));
            }
        } else {
            for (int c = lane; c < V; c += 32) row[c] = 0.0f;
        }
        __syncwarp();
        if (lane == 0) {
            row[0] -= g * o.x;
            if (u < W - 1) row[y[u]] -= g * o.y;                 // both land on class 0 if y[u] == 0
        }
        float* dst = p.gx + (ubase + (size_t)t * p.U1 + u) * V;
        if (p.use_bulk) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(dst, row, (uint32_t)V * 4u); bulk_commit(); }
        } else {
            __syncwarp();
            for (int c = lane; c < V; c += 32) dst[c] = row[c];
        }
    }
    if (p.use_bulk && lane == 0) bulk_wait_all<0>();
}

// grid (ceil(T*U1/64), N), block 256: zero the gradient rows of padded nodes
__global__ void __launch_bounds__(256) rnnt_zero_kernel(RnntGradParams p) {
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x, W = mt.y + 1;
    if (Tn == p.T && W == p.U1) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 64;
    for (int r = r0 + warp; r < min(r0 + 64, p.T * p.U1); r += 8) {
        const int t = r / p.U1, u = r - t * p.U1;
        if (t < Tn && u < W) continue;
        float* dst = p.gx + ((size_t)n * p.T * p.U1 + r) * p.V;
        if ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = lane; c < (p.V >> 2); c += 32) ((float4*)dst)[c] = z;
        } else {
            for (int c = lane; c < p.V; c += 32) dst[c] = 0.0f;
        }
    }
}

}  // namespace hab

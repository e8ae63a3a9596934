// rnnt.cuh — RNN-T lattice loss + joint-logit gradient for sm_100a.  Replaces
// ha/transducer.py:175-205 (transducer_forward_score; its column scans ha/scan.py:88-126 become an
// anti-diagonal sweep) and the autograd backward.
//
//   rnnt_prep_kernel     lengths/targets -> int32 metadata
//   rnnt_rows_kernel     one warp per lattice node (n,t,u): row log-sum-exp + the blank and label
//                        log-probs, written in a diagonal-major ("skewed") layout
//   rnnt_lattice_kernel  one CTA per utterance, one thread per u: alpha and beta swept along the
//                        anti-diagonals concurrently (two halves of the CTA) in the linear domain on
//                        extended-range numbers (fp32 mantissa + int32 exponent), then one parallel
//                        arc-occupancy pass
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
    w.bl = take(sizeof(float2) * (size_t)N * w.D * U1);         // skewed (t+u, u); probability = x * 2^(int)y
    w.lb = take(sizeof(float2) * (size_t)N * w.D * U1);
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
    const float* x;            // (N,T,U1,V) view: element strides sx_n, sx_t, sx_u, unit class stride
    long long sx_n, sx_t, sx_u;
    int N, T, U1, V;
    const int4* meta; const int* tgt; int Up;
    float* lse2; float2* bl; float2* lb; int D;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

// 2^x for an fp32 log2-probability x = pm * 2^K, pm in [0.707, 1.415], as (pm, bits of K)
__device__ __forceinline__ float2 log2_to_parts(float x) {
    const bool tiny = !(x > -1.0e6f);
    const float k = rintf(tiny ? 0.0f : x);
    return make_float2(tiny ? 1.0f : exp2_poly(x - k), __int_as_float(tiny ? -(1 << 24) : (int)k));
}

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
    const float* xb = p.x + (long long)n * p.sx_n;
    const int* y = p.tgt + (size_t)n * p.Up;

    auto issue = [&](int r) {
        const int stage = r % nstage;
        const int v = v0 + warp + nw * r;
        const int t = v / W, u = v - t * W;
        const float* src = xb + (long long)t * p.sx_t + (long long)u * p.sx_u;
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
        if (lane < 2) {     // lane 0: blank, lane 1: label
            const size_t sk = ((size_t)n * p.D + (t + u)) * p.U1 + u;
            if (lane == 0) p.lse2[((size_t)n * p.T + t) * p.U1 + u] = l2;
            const float v = lane ? ((u < W - 1) ? fmaf(row[y[u < W - 1 ? u : 0]], kLog2e, -l2) : kVoid) : fmaf(row[0], kLog2e, -l2);
            (lane ? p.lb : p.bl)[sk] = log2_to_parts(v);
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- lattice ---
struct RnntLatticeParams {
    int N, T, U1, D;
    const int4* meta;
    const float2* bl; const float2* lb;
    double* alpha; double* beta; float2* occ;
    float* loss; float* loss_ws;
};

// Lattice values are extended-range linear numbers (common.cuh, XF: fp32 mantissa in [1, 2) + int32
// exponent).  alpha/beta are sums of products of probabilities, so the recursion is one aligned FADD and
// one FMUL per arc (no log-add-exp); every rounding is 6e-8 relative whatever the magnitude.

__device__ __forceinline__ XF xf_mul_parts(XF a, float2 p) {       // p = (mantissa, exponent bits), rnnt_rows_kernel
    XF r = xf_mul_norm(a, p.x);
    r.e += __float_as_int(p.y);
    return r;
}
__device__ __forceinline__ XF xf_normalise(XF a) {
    const int bits = __float_as_int(a.m);
    return xf_make(__int_as_float((bits & 0x007fffff) | 0x3f800000), a.e + (bits >> 23) - 127);
}

// grid N, block sides * round_up(U1, 32).  Thread u of a side owns column u; on anti-diagonal d it is
// at t = d - u.  With sides == 2 the alpha sweep (threads [0, half)) and the beta sweep (threads
// [half, 2 half)) run concurrently, each behind its own named barrier; with sides == 1 (U+1 > 512)
// the same threads run them one after the other.  Arc occupancies are then one parallel pass.
//   alpha[t,u] = alpha[t-1,u] blank[t-1,u] + alpha[t,u-1] label[t,u-1]        ha/transducer.py:197-202
// Each thread forms the two outgoing arc products of its node right away (the one along u travels to
// lane u+1 by shuffle), and reads and decodes its node's emissions kPre diagonals ahead.
constexpr int kPre = 8;
template <int MAXT>
__global__ void __launch_bounds__(MAXT) rnnt_lattice_kernel(RnntLatticeParams p) {
    __shared__ int2 s_edge[2][2][32];       // [side][diagonal parity][warp]
    __shared__ int2 s_z;
    const int n = blockIdx.x;
    const int U1 = p.U1;
    const int half = round_up(U1, 32);
    const int sides = blockDim.x / half;
    const int side = threadIdx.x / half;
    const int u = threadIdx.x - side * half, lane = u & 31, warp = u >> 5, nwarp = half >> 5;
    const int4 mt = p.meta[n];
    const int Tn = mt.x, Un = mt.y;
    if (mt.z) { if (threadIdx.x == 0) { p.loss[n] = CUDART_NAN_F; p.loss_ws[n] = CUDART_NAN_F; } return; }
    const float2* __restrict__ bl = p.bl + (size_t)n * p.D * U1;
    const float2* __restrict__ lb = p.lb + (size_t)n * p.D * U1;
    int2* __restrict__ al = (int2*)(p.alpha + (size_t)n * p.D * U1);
    int2* __restrict__ be = (int2*)(p.beta + (size_t)n * p.D * U1);
    float2* __restrict__ occ = p.occ + (size_t)n * p.D * U1;
    const int nd = Tn + Un;                 // diagonals 0 .. nd-1
    const bool col = u <= Un;
    const XF vd = xf_make(1.0f, kVoidE);
    const int2 vd2 = make_int2(__float_as_int(1.0f), kVoidE);
    const float2 one2 = make_float2(1.0f, 0.0f);
    auto pack = [](XF a) { return make_int2(__float_as_int(a.m), a.e); };
    // Both sweeps run whole blocks of kPre diagonals: the steps past the last diagonal are no-ops (no node
    // is valid there), which keeps the per-step code free of bounds branches.

    if (side == 0) {
        // ---- alpha, forward over diagonals
        XF ob = vd, ol = vd;                // my previous node times its blank / label emission
        float2 fb[kPre], fl[kPre];          // emissions of my node on diagonals d .. d+kPre-1
        const float2* pb = bl + u;          // -> diagonal d + kPre
        const float2* pl = lb + u;
        int2* pa = al + u;                  // -> diagonal d
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const bool in = col && k < nd;
            fb[k] = in ? pb[0] : one2;
            fl[k] = in ? pl[0] : one2;
            pb += U1; pl += U1;
        }
        const int2* ein = &s_edge[0][0][warp > 0 ? warp - 1 : 0];
        int2* eout = &s_edge[0][0][warp];
        for (int d0 = 0; d0 < nd; d0 += kPre) {
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int d = d0 + k;
                const float2 cb = fb[k], cl = fl[k];
                {
                    const bool in = col && d + kPre < nd;
                    fb[k] = in ? pb[0] : one2;
                    fl[k] = in ? pl[0] : one2;
                    pb += U1; pl += U1;
                }
                float xm = __shfl_up_sync(0xffffffffu, ol.m, 1);           // alpha[t, u-1] label[t, u-1]
                int xe = __shfl_up_sync(0xffffffffu, ol.e, 1);
                {
                    const int2 in = ein[((d + 1) & 1) * 32];
                    if (lane == 0) { xm = (warp > 0) ? __int_as_float(in.x) : 1.0f; xe = (warp > 0) ? in.y : kVoidE; }
                }
                const bool valid = col && (unsigned)(d - u) < (unsigned)Tn;
                XF cur = (d == 0) ? xf_make(1.0f, 0) : xf_normalise(xf_add(ob, xf_make(xm, xe)));   // alpha[0,0] = 1
                if (valid) *pa = pack(cur);
                pa += U1;
                ob = valid ? xf_mul_parts(cur, cb) : vd;
                ol = (valid && u < Un) ? xf_mul_parts(cur, cl) : vd;
                if (lane == 31) eout[(d & 1) * 32] = pack(ol);
                // Z = alpha[T-1,U] blank[T-1,U]                                  ha/transducer.py:204-205
                if (d == nd - 1 && u == Un) s_z = pack(ob);
                named_bar_sync(1, half);
            }
        }
    }
    if (side == sides - 1) {
        // ---- beta, backward over diagonals
        XF b = vd;                          // beta[t+1,u]: my node on the next diagonal
        float2 fb[kPre], fl[kPre];
        const float2* pb = bl + (size_t)(nd - 1) * U1 + u;     // -> diagonal d - kPre
        const float2* pl = lb + (size_t)(nd - 1) * U1 + u;
        int2* pe = be + (size_t)(nd - 1) * U1 + u;             // -> diagonal d
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const bool in = col && nd - 1 - k >= 0;
            fb[k] = in ? pb[0] : one2;
            fl[k] = in ? pl[0] : one2;
            pb -= U1; pl -= U1;
        }
        const int2* ein = &s_edge[1][0][warp + 1 < nwarp ? warp + 1 : warp];
        int2* eout = &s_edge[1][0][warp];
        for (int d0 = nd - 1; d0 >= 0; d0 -= kPre) {
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int d = d0 - k;
                const int t = d - u;
                const float2 cb = fb[k], cl = fl[k];
                {
                    const bool in = col && d - kPre >= 0;
                    fb[k] = in ? pb[0] : one2;
                    fl[k] = in ? pl[0] : one2;
                    pb -= U1; pl -= U1;
                }
                float xm = __shfl_down_sync(0xffffffffu, b.m, 1);          // beta[t, u+1]
                int xe = __shfl_down_sync(0xffffffffu, b.e, 1);
                {
                    const int2 in = ein[((d + 1) & 1) * 32];
                    if (lane == 31) { xm = (warp + 1 < nwarp) ? __int_as_float(in.x) : 1.0f; xe = (warp + 1 < nwarp) ? in.y : kVoidE; }
                }
                const bool valid = col && (unsigned)t < (unsigned)Tn;
                // only the terminal blank leaves the last frame
                const bool lastf = t == Tn - 1;
                const XF bsrc = lastf ? xf_make(1.0f, (u == Un) ? 0 : kVoidE) : b;
                const XF tb = xf_mul_parts(bsrc, cb);
                const XF tl = xf_mul_parts(xf_make(xm, (u < Un) ? xe : kVoidE), cl);
                const XF sum = xf_normalise(xf_add(tb, tl));
                const XF cur = valid ? sum : vd;
                if (valid) *pe = pack(cur);
                pe -= U1;
                b = cur;
                if (lane == 0) eout[(d & 1) * 32] = pack(b);
                named_bar_sync(2, half);
            }
        }
    }
    __syncthreads();
    const float zm = __int_as_float(s_z.x);
    const int ze = s_z.y;
    const bool feasible = ze > kVoidETest;
    if (threadIdx.x == 0) {
        const float v = feasible ? (float)(-((double)ze + log2((double)zm)) * kLn2) : CUDART_INF_F;
        p.loss[n] = v; p.loss_ws[n] = v;
    }
    // ---- arc occupancies: occ_blank = alpha blank beta[t+1,u] / Z, occ_label = alpha label beta[t,u+1] / Z
    const float rz = 1.0f / zm;
    auto arc = [&](int2 a, int2 b2, float2 x) {
        const float m = __int_as_float(a.x) * __int_as_float(b2.x) * x.x * rz;         // < 6
        const int ex = min(max(a.y + b2.y + __float_as_int(x.y) - ze, -127), 2);
        return m * __int_as_float((ex + 127) << 23);
    };
    // four nodes per thread and pass: all loads first, so that their latencies overlap
    const int total = nd * U1;
    for (int k0 = threadIdx.x; k0 < total; k0 += 4 * blockDim.x) {
        int2 a[4], b0[4], b1[4]; float2 xb[4], xl[4]; int flag[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + q * blockDim.x;
            const int d = k / U1, uu = k - d * U1, t = d - uu;
            const bool in = k < total && uu <= Un && t >= 0 && t < Tn;
            const bool hb = in && t < Tn - 1, hl = in && uu < Un;
            flag[q] = (in ? 1 : 0) | (hb ? 2 : 0) | (hl ? 4 : 0) | ((in && t == Tn - 1 && uu == Un) ? 8 : 0);
            const int2 one = make_int2(__float_as_int(1.0f), 0);
            a[q] = in ? al[k] : one;
            xb[q] = in ? bl[k] : make_float2(1.0f, 0.0f);
            xl[q] = hl ? lb[k] : make_float2(1.0f, 0.0f);
            b0[q] = hb ? be[(size_t)(d + 1) * U1 + uu] : one;
            b1[q] = hl ? be[(size_t)(d + 1) * U1 + uu + 1] : one;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!(flag[q] & 1)) continue;
            float2 o = make_float2(0.0f, 0.0f);
            if (feasible) {
                if (flag[q] & (2 | 8)) o.x = arc(a[q], b0[q], xb[q]);     // terminal blank: b0 = 1
                if (flag[q] & 4) o.y = arc(a[q], b1[q], xl[q]);
            }
            occ[k0 + q * blockDim.x] = o;
        }
    }
}

// ------------------------------------------------------------------------------------ grad ---
struct RnntGradParams {
    const float* x; float* gx;
    long long sx_n, sx_t, sx_u, sg_n, sg_t, sg_u;      // element strides of the joint view and of its gradient
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
        const float* src = p.x + (long long)n * p.sx_n + (long long)t * p.sx_t + (long long)u * p.sx_u;
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
        float* dst = p.gx + (long long)n * p.sg_n + (long long)t * p.sg_t + (long long)u * p.sg_u;
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
        float* dst = p.gx + (long long)n * p.sg_n + (long long)t * p.sg_t + (long long)u * p.sg_u;
        if ((p.V & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c = lane; c < (p.V >> 2); c += 32) ((float4*)dst)[c] = z;
        } else {
            for (int c = lane; c < p.V; c += 32) dst[c] = 0.0f;
        }
    }
}

}  // namespace hab

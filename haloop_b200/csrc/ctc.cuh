// ctc.cuh — CTC loss + logit gradient for sm_100a.  Replaces ha/ctc.py:110-174
// (ctc_forward_score3) and its autograd backward.  Three kernels + a tiny prep:
//
//   ctc_prep_kernel     targets/lengths -> int32 metadata, duplicate-label chains, work order
//   ctc_rows_kernel     one warp per (b,t) row: log-softmax statistics + gather of the blank and
//                       label emissions (rows staged in shared memory by bulk async copies)
//   ctc_trellis_kernel  alpha and beta recursions as two warps per utterance that meet in the
//                       middle; extended-label states live in registers as (blank,label) pairs
//                       interleaved across the 32 lanes; emission rows and the other side's stored
//                       trellis rows are prefetched through a TMA (cp.async.bulk) ring; posterior
//                       occupancies overwrite the emission buffer in place
//   ctc_grad_kernel     one warp per row: softmax - occupancy, times grad_out, written once
//
// All log-domain values are log2.  alpha/beta are kept relative to per-slot (64-state) integer
// offsets so fp32 state stays O(100) in magnitude whatever T is (SURVEY.md finding 3).
#pragma once
#include "common.cuh"

namespace hab {

struct CtcWs {                 // workspace layout (byte offsets), filled by ctc_ws_layout()
    size_t meta, order, tgt, dupnext, loss, lse2, em, tr, total;
    int Sp, E, JWp, SPX;
};

__host__ inline CtcWs ctc_ws_layout(int T, int N, int S) {
    CtcWs w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.E = round_up(S + 1, 4);                       // blank + S labels per emission row
    w.JWp = round_up((S + 1 + 31) / 32, 4);         // per-slot offsets stored in front of a trellis row
    w.SPX = w.JWp + round_up(2 * S + 2, 4);         // + (blank,label) pairs
    size_t o = 256;                                 // header
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.dupnext = take(sizeof(int) * (size_t)N * w.Sp);
    w.loss = take(sizeof(float) * (size_t)N);       // copy of the per-utterance loss for the backward pass
    w.lse2 = take(sizeof(float) * (size_t)N * T);
    w.em = take(sizeof(float) * (size_t)N * T * w.E);
    w.tr = take(sizeof(float) * (size_t)N * T * w.SPX);
    w.total = o;
    return w;
}

// ------------------------------------------------------------------------------------ prep ---
struct PrepParams {
    const void* targets; long long tgt_stride; int tgt64;
    const void* in_len; const void* tgt_len; int len64;
    int T, N, V, S, Sp;
    int4* meta; int* order; int* tgt; int* dupnext;
    int star;   // duplicate chains also cover position L_n (< S): star-CTC's last star reads targets[n, L_n]
};

// grid N, block 128.  meta[n] = {T_n, L_n, invalid, 0}.
__global__ void __launch_bounds__(128) ctc_prep_kernel(PrepParams p) {
    extern __shared__ int s_y[];
    __shared__ int s_bad, s_rank;
    const int n = blockIdx.x;
    long long Tn = load_idx(p.in_len, n, p.len64), Ln = load_idx(p.tgt_len, n, p.len64);
    if (threadIdx.x == 0) { s_bad = (Tn < 0 || Tn > p.T || Ln < 0 || Ln > p.S) ? 1 : 0; s_rank = 0; }
    __syncthreads();
    const int L = s_bad ? 0 : (int)Ln;
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) {
        long long y = load_idx(p.targets, (long long)n * p.tgt_stride + k, p.tgt64);
        // labels beyond L_n are kept too: star-CTC reads targets[n, L_n] (ha/star.py:46)
        if (y < 0 || y >= p.V) { if (k < L) s_bad = 1; y = 0; }
        s_y[k] = (int)y;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) {
        int y = s_y[k], nxt = -1, notfirst = 0;
        const int Lc = p.star ? min(L + 1, p.S) : L;
        if (k < Lc) {
            for (int j = k + 1; j < Lc; ++j) if (s_y[j] == y) { nxt = j; break; }
            for (int j = 0; j < k; ++j) if (s_y[j] == y) { notfirst = 1; break; }
        }
        p.tgt[(size_t)n * p.Sp + k] = y | (notfirst ? kNotFirst : 0);
        p.dupnext[(size_t)n * p.Sp + k] = nxt;
    }
    __syncthreads();
    {   // longest-first work order (rank by counting; N is a batch size)
        const bool lenbad = (Tn < 0 || Tn > p.T || Ln < 0 || Ln > p.S);
        const long long mine = lenbad ? 0 : Tn * (Ln + 1);
        int rank = 0;
        for (int m = threadIdx.x; m < p.N; m += blockDim.x) {
            long long Tm = load_idx(p.in_len, m, p.len64), Lm = load_idx(p.tgt_len, m, p.len64);
            long long c = (Tm < 0 || Tm > p.T || Lm < 0 || Lm > p.S) ? 0 : Tm * (Lm + 1);
            rank += (c > mine) || (c == mine && m < n);
        }
        if (rank) atomicAdd(&s_rank, rank);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.meta[n] = make_int4(s_bad ? 0 : (int)Tn, L, s_bad, 0);
        p.order[s_rank] = n;
    }
}

// ------------------------------------------------------------------------------------ rows ---
struct RowsParams {
    const float* x; long long sx_t, sx_n;
    int T, N, V;
    const int4* meta; const int* tgt; int Sp;
    float* lse2; float* em; int E;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

constexpr int kMaxRowWarps = 8;

__host__ __device__ inline size_t rows_smem_bytes(int Sp, int V, int nstage, int nwarps) {
    // [mbarriers][targets][row ring per warp]
    return round_up_sz((size_t)nwarps * nstage * 8, 128) + round_up_sz((size_t)Sp * 4, 128) +
           (size_t)nwarps * nstage * V * 4;
}

// grid (ceil(T / (nwarps*rows_per_warp)), N), block 32*nwarps.  Warp w takes rows t0+w, t0+w+nwarps, ...
template <bool VEC4>
__global__ void __launch_bounds__(kMaxRowWarps * 32) ctc_rows_kernel(RowsParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kRowWarps = p.nwarps;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const int Tn = mt.z ? 0 : mt.x, L = mt.y;
    const int t0 = blockIdx.x * (kRowWarps * p.rows_per_warp);
    if (t0 >= Tn) return;
    const int V = p.V, nstage = p.nstage;
    uint64_t* s_bar = (uint64_t*)smem_raw;
    int* s_tgt = (int*)(smem_raw + round_up_sz((size_t)kRowWarps * nstage * 8, 128));
    float* s_rows = (float*)((unsigned char*)s_tgt + round_up_sz((size_t)p.Sp * 4, 128));
    for (int k = threadIdx.x; k < L; k += blockDim.x) s_tgt[k] = p.tgt[(size_t)n * p.Sp + k] & kLabelMask;
    float* wrows = s_rows + (size_t)warp * nstage * V;
    uint64_t* wbar = s_bar + warp * nstage;
    if (lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&wbar[s], 1);
    mbar_init_fence();
    __syncthreads();

    int nrows = 0;
    if (t0 + warp < Tn) nrows = min(p.rows_per_warp, (Tn - 1 - t0 - warp) / kRowWarps + 1);
    const float* xb = p.x + (long long)n * p.sx_n;

    auto issue = [&](int r) {
        const int stage = r % nstage;
        const float* src = xb + (long long)(t0 + warp + kRowWarps * r) * p.sx_t;
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
        const int t = t0 + warp + kRowWarps * r;
        float l2 = 0.0f;                       // log2-sum-exp2 of the row; 0 when x already holds log-probs
        if (p.from_logits) {
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
            float s = 0.0f;
            if (VEC4) {
                const float4* r4 = (const float4*)row;
                for (int c = lane; c < (V >> 2); c += 32) {
                    float4 v = r4[c];
                    s += ex2f(fmaf(v.x, kLog2e, -m2)) + ex2f(fmaf(v.y, kLog2e, -m2)) +
                         ex2f(fmaf(v.z, kLog2e, -m2)) + ex2f(fmaf(v.w, kLog2e, -m2));
                }
            } else {
                for (int c = lane; c < V; c += 32) s += ex2f(fmaf(row[c], kLog2e, -m2));
            }
            s = warp_sum(s);
            l2 = m2 + log2f(s);
        }
        float* erow = p.em + ((size_t)n * p.T + t) * p.E;
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            erow[0] = fmaf(row[0], kLog2e, -l2);
        }
        for (int k = lane; k < L; k += 32) erow[1 + k] = fmaf(row[s_tgt[k]], kLog2e, -l2);
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- trellis ---
struct TrellisParams {
    int T, N;
    const int4* meta; const int* order; const int* tgt; int Sp;
    float* em; int E;          // emissions in, occupancies out (in place)
    float* tr; int SPX, JWp;   // row = [JWp slot offsets (int)] [2*(L+1) floats: (blank,label) pairs]
    float* loss; float* loss_ws;
    int nstage; int warp_bytes;
};

__host__ __device__ inline int trellis_warp_bytes(int E, int SPX, int nstage) {
    return round_up((nstage * (E + SPX) + 2 * E) * 4 + 2 * nstage * 8, 128);
}

constexpr int kRenorm = 4;   // steps between per-slot renormalisations

// grid ceil(N/2), block 128: warps (2u, 2u+1) are the alpha and beta side of one utterance.
// Side d walks time from its own end: step i is frame t = d ? T-1-i : i, on its own ordering of
// the label pairs (beta = alpha on the reversed label sequence).  Phase 1 (first half of the
// frames) stores every trellis row; phase 2 combines live rows with the rows the other side
// stored, so posteriors need T sequential steps instead of 2T and nothing is recomputed.
template <int J>
__global__ void __launch_bounds__(128, 1) ctc_trellis_kernel(TrellisParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int usel = warp >> 1, dir = warp & 1;
    const int idx = blockIdx.x * 2 + usel;
    if (idx >= p.N) return;
    const int n = p.order[idx];
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    if (mt.z || Tn == 0) {
        const float v = mt.z ? CUDART_NAN_F : ((L == 0) ? 0.0f : CUDART_INF_F);
        if (dir == 0 && lane == 0) { p.loss[n] = v; p.loss_ws[n] = v; }
        return;
    }
    const int P = L + 1;
    const int nslot = (P + 31) >> 5;
    const int nstage = p.nstage, E = p.E, SPX = p.SPX, JWp = p.JWp;

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

    // skip transitions into my label states (ha/ctc.py:140-142), in my own direction
    unsigned allowed = 0;
    {
        const int* y = p.tgt + (size_t)n * p.Sp;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int q = 32 * j + lane;
            if (q >= 1 && q < L) {
                const int ycur = y[dir ? L - 1 - q : q] & kLabelMask;
                const int yprv = y[dir ? L - q : q - 1] & kLabelMask;
                const int dst = dir ? yprv : ycur;     // the label entered in forward time
                if (dst != 0 && ycur != yprv) allowed |= 1u << j;
            }
        }
    }

    float a0[J], a1[J];
    int off[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { a0[j] = kVoid; a1[j] = kVoid; off[j] = 0; }

    float* em_base = p.em + (size_t)n * p.T * E;
    float* tr_base = p.tr + (size_t)n * p.T * SPX;
    const uint32_t em_bytes = (uint32_t)round_up(P, 4) * 4u;
    const uint32_t tr_bytes = (uint32_t)(JWp + round_up(2 * P, 4)) * 4u;
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;

    auto issue_em = [&](int i) {          // lane 0 only
        const int st = i % nstage, t = dir ? Tn - 1 - i : i;
        mbar_expect_tx(&bar_em[st], em_bytes);
        bulk_g2s(em_ring + st * E, em_base + (size_t)t * E, em_bytes, &bar_em[st]);
    };
    auto issue_tr = [&](int k) {          // lane 0 only; k-th phase-2 step
        const int st = k % nstage, i = steps1 + k, t = dir ? Tn - 1 - i : i;
        mbar_expect_tx(&bar_tr[st], tr_bytes);
        bulk_g2s(tr_ring + st * SPX, tr_base + (size_t)t * SPX, tr_bytes, &bar_tr[st]);
    };
    auto phase_switch = [&]() {
        // my stored rows -> visible to the sibling warp's bulk (async-proxy) loads, and vice versa
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
        const int st = i % nstage;
        const int t = dir ? Tn - 1 - i : i;
        mbar_wait(&bar_em[st], (uint32_t)(i / nstage) & 1u);
        const float* er = em_ring + st * E;
        const float eb = er[0];
        float el[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int q = 32 * j + lane;
            el[j] = (q < L) ? er[1 + (dir ? L - 1 - q : q)] : kVoid;
        }
        __syncwarp();
        if (lane == 0 && i + nstage < Tn) issue_em(i + nstage);

        if (i == 0) {
            if (lane == 0) { a0[0] = eb; a1[0] = el[0]; }     // ha/ctc.py:138 (el is void when L == 0)
        } else {
            float c[J];
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (j < nslot) c[j] = __shfl_sync(0xffffffffu, a1[j], (lane + 31) & 31);
            if (lane == 0) {
#pragma unroll
                for (int j = J - 1; j >= 1; --j)
                    if (j < nslot) c[j] = c[j - 1] + (float)(off[j - 1] - off[j]);
                c[0] = kVoid;
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const float u = lae2(a0[j], c[j]);            // self (+) previous label -> blank
                    const float sel = ((allowed >> j) & 1u) ? u : a0[j];
                    a1[j] = fmaxf(lae2(sel, a1[j]) + el[j], kVoid);
                    a0[j] = fmaxf(u + eb, kVoid);
                }
            }
        }
        if ((i % kRenorm) == kRenorm - 1) {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const float m = warp_max(fmaxf(a0[j], a1[j]));
                    if (m > kVoidTest) {
                        const float k = rintf(m);
                        a0[j] -= k; a1[j] -= k; off[j] += (int)k;
                    }
                }
            }
        }
        if (i < steps1) {
            float* row = tr_base + (size_t)t * SPX;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const int q = 32 * j + lane;
                    if (q < P) ((float2*)(row + JWp))[q] = make_float2(a0[j], a1[j]);
                    if (lane == j) ((int*)row)[j] = off[j];
                }
            }
        } else {
            const int k = i - steps1;
            const int ts = k % nstage;
            mbar_wait(&bar_tr[ts], (uint32_t)(k / nstage) & 1u);
            const float* orow = tr_ring + ts * SPX + JWp;
            const int* ooff = (const int*)(tr_ring + ts * SPX);
            // the other side's state for my state s is its state 2L - s
            float v0[J], v1[J];
            int i0[J], i1[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                v0[j] = kVoid; v1[j] = kVoid; i0[j] = 0; i1[j] = 0;
                if (j < nslot) {
                    const int q = 32 * j + lane;
                    if (q < P) {
                        const int r0 = 2 * (L - q);
                        v0[j] = a0[j] + orow[r0] - eb;
                        i0[j] = off[j] + ooff[r0 >> 6];
                    }
                    if (q < L) {
                        const int r1 = 2 * (L - q) - 1;
                        v1[j] = a1[j] + orow[r1] - el[j];
                        i1[j] = off[j] + ooff[r1 >> 6];
                    }
                }
            }
            if (k == 0) {
                double mx = -1.0e300;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (j < nslot) mx = fmax(mx, fmax((double)i0[j] + (double)v0[j], (double)i1[j] + (double)v1[j]));
                mx = warp_max_d(mx);
                feasible = mx > (double)kVoidTest;
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if (j < nslot)
                        s += ex2f((float)((double)i0[j] + (double)v0[j] - mx)) +
                             ex2f((float)((double)i1[j] + (double)v1[j] - mx));
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
            if (lane == 0) bulk_wait_read<1>();     // the store issued two steps ago has left this buffer
            __syncwarp();
            float bsum = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nslot) {
                    const int q = 32 * j + lane;
                    const float g0 = feasible ? ex2f(v0[j] + (float)(i0[j] - IZ) - fZ) : 0.0f;
                    const float g1 = feasible ? ex2f(v1[j] + (float)(i1[j] - IZ) - fZ) : 0.0f;
                    bsum += g0;
                    if (q < L) ob[1 + (dir ? L - 1 - q : q)] = g1;
                }
            }
            bsum = warp_sum(bsum);
            if (lane == 0) ob[0] = bsum;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(em_base + (size_t)t * E, ob, em_bytes);
                bulk_commit();
                if (k + nstage < Tn - steps1) issue_tr(k + nstage);
            }
        }
    }
    if (steps1 == Tn) phase_switch();     // only T == 1, beta side: still owes the barrier
    if (lane == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------------------------ grad ---
struct GradParams {
    const float* x; long long sx_t, sx_n;
    float* gx; long long sg_t, sg_n;
    int T, N, V;
    const int4* meta; const int* tgt; const int* dupnext; int Sp;
    const float* lse2; const float* occ; int E;
    const float* gout; const float* loss;
    int from_logits, use_bulk, rows_per_warp, nstage, nwarps;
};

__host__ __device__ inline size_t grad_smem_bytes(int Sp, int V, int E, int nstage, int nwarps) {
    // [mbarriers][targets + duplicate chains][(x row, occupancy row) ring per warp]
    return round_up_sz((size_t)nwarps * nstage * 8, 128) + round_up_sz((size_t)Sp * 8, 128) +
           (size_t)nwarps * nstage * (V + E) * 4;
}

// grid (ceil(T / (nwarps*rows_per_warp)), N), block 32*nwarps.  d loss / d logits = (softmax - occupancy) * gout
// (from_logits) or -occupancy * gout (log-prob input, the reference's autograd boundary); rows at
// t >= T_n, and every row of an infeasible or invalid utterance, are zero.
template <bool VEC4>
__global__ void __launch_bounds__(kMaxRowWarps * 32) ctc_grad_kernel(GradParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kRowWarps = p.nwarps;
    const int n = blockIdx.y;
    const int4 mt = p.meta[n];
    const int L = mt.y;
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x;     // NaN/inf loss -> all-zero gradient
    const int t0 = blockIdx.x * (kRowWarps * p.rows_per_warp);
    const int V = p.V, E = p.E, nstage = p.nstage;
    uint64_t* s_bar = (uint64_t*)smem_raw;
    int* s_tgt = (int*)(smem_raw + round_up_sz((size_t)kRowWarps * nstage * 8, 128));
    int* s_nxt = s_tgt + p.Sp;
    float* s_rows = (float*)((unsigned char*)s_tgt + round_up_sz((size_t)p.Sp * 8, 128));
    if (t0 < Tn) {
        for (int k = threadIdx.x; k < L; k += blockDim.x) {
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

    // rows t0+warp+8r < min(T, t0+8*rpw); the first `nreal` of them are inside the utterance
    int nall = 0, nreal = 0;
    if (t0 + warp < p.T) nall = min(p.rows_per_warp, (p.T - 1 - t0 - warp) / kRowWarps + 1);
    if (t0 + warp < Tn) nreal = min(nall, (Tn - 1 - t0 - warp) / kRowWarps + 1);
    const float* xb = p.x + (long long)n * p.sx_n;
    float* gb = p.gx + (long long)n * p.sg_n;
    const float g = p.gout[n];
    const uint32_t occ_bytes = (uint32_t)round_up(L + 1, 4) * 4u;

    auto issue = [&](int r) {
        const int stage = r % nstage;
        const int t = t0 + warp + kRowWarps * r;
        float* dst = wrows + (size_t)stage * (V + E);
        const float* osrc = p.occ + ((size_t)n * p.T + t) * E;
        if (p.use_bulk) {
            if (lane == 0) {
                const uint32_t xbytes = p.from_logits ? (uint32_t)V * 4u : 0u;
                mbar_expect_tx(&wbar[stage], xbytes + occ_bytes);
                if (p.from_logits) bulk_g2s(dst, xb + (long long)t * p.sx_t, xbytes, &wbar[stage]);
                bulk_g2s(dst + V, osrc, occ_bytes, &wbar[stage]);
            }
        } else {
            if (p.from_logits) {
                const float* src = xb + (long long)t * p.sx_t;
                for (int c = lane; c < V; c += 32) dst[c] = src[c];
            }
            for (int c = lane; c <= L; c += 32) dst[V + c] = osrc[c];
        }
    };
    for (int r = 0; r < min(nstage - 1, nreal); ++r) issue(r);

    for (int r = 0; r < nreal; ++r) {
        const int stage = r % nstage;
        // keep nstage-1 loads in flight; the stage being refilled was stored from one row ago
        if (r + nstage - 1 < nreal) {
            if (p.use_bulk && lane == 0) bulk_wait_read<0>();
            __syncwarp();
            issue(r + nstage - 1);
        }
        if (p.use_bulk) mbar_wait(&wbar[stage], (uint32_t)(r / nstage) & 1u);
        else __syncwarp();
        float* row = wrows + (size_t)stage * (V + E);
        const float* occ = row + V;
        const int t = t0 + warp + kRowWarps * r;
        if (p.from_logits) {
            const float l2 = p.lse2[(size_t)n * p.T + t];
            if (VEC4) {
                float4* r4 = (float4*)row;
                for (int c = lane; c < (V >> 2); c += 32) {
                    float4 v = r4[c];
                    v.x = g * ex2f(fmaf(v.x, kLog2e, -l2)); v.y = g * ex2f(fmaf(v.y, kLog2e, -l2));
                    v.z = g * ex2f(fmaf(v.z, kLog2e, -l2)); v.w = g * ex2f(fmaf(v.w, kLog2e, -l2));
                    r4[c] = v;
                }
            } else {
                for (int c = lane; c < V; c += 32) row[c] = g * ex2f(fmaf(row[c], kLog2e, -l2));
            }
        } else {
            for (int c = lane; c < V; c += 32) row[c] = 0.0f;
        }
        __syncwarp();
        // occupancy of class y = sum over the positions that carry y, walked in position order by
        // the first such position: one writer per class, no atomics, deterministic
        for (int k = lane; k < L; k += 32) {
            const int w = s_tgt[k];
            if (!(w & kNotFirst)) {
                float s = occ[1 + k];
                for (int j = s_nxt[k]; j >= 0; j = s_nxt[j]) s += occ[1 + j];
                row[w & kLabelMask] -= g * s;
            }
        }
        __syncwarp();
        if (lane == 0) row[0] -= g * occ[0];
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
        float* dstg = gb + (long long)(t0 + warp + kRowWarps * r) * p.sg_t;
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

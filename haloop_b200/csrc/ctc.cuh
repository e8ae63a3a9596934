// ctc.cuh — CTC loss + logit gradient for sm_100a.  Replaces ha/ctc.py:110-174
// (ctc_forward_score3) and its autograd backward.  Three kernels + a tiny prep:
//
//   ctc_prep_kernel     targets/lengths -> int32 metadata, duplicate-label chains, work order
//   ctc_rows_kernel     one warp per (b,t) row: log-softmax statistics + gather of the blank and
//                       label emissions (rows staged in shared memory by bulk async copies)
//   ctc_trellis_kernel  one CTA per utterance: W warps sweep alpha forward and W warps sweep beta
//                       backward until they meet in the middle; extended-label states live in
//                       registers as (blank,label) pairs across the lanes, neighbours exchanged by
//                       warp shuffles (and a 2-float mailbox between warps); emission rows and the
//                       other side's stored trellis rows are prefetched through a TMA
//                       (cp.async.bulk) ring; posterior occupancies overwrite the emission buffer
//   ctc_grad_kernel     one warp per row: softmax - occupancy, times grad_out, written once
//
// alpha/beta are extended-range linear numbers (fp32 mantissa + int32 exponent, common.cuh XF) and the
// gathered emissions are linear fp32 probabilities relative to a per-row integer shift, so every fp32
// rounding in the recursions is 6e-8 relative whatever T and log V are (SURVEY.md finding 3).
#pragma once
#include "common.cuh"

namespace hab {

struct CtcWs {                 // workspace layout (byte offsets), filled by ctc_ws_layout()
    size_t meta, order, tgt, dupnext, loss, lse2, em, tr, total;
    int Sp, E, JWp, SPX;       // E: floats per emission row (4 header floats + one per label)
};

// CTC emission row (floats): [0] ct (integer row shift)  [1] blank  [4+k] label k, where an emission is
// the probability 2^(log2 p - ct) as a plain fp32 (<= 2^0.5, clamped below at 2^-126): the trellis runs in
// the linear domain on extended-range numbers (common.cuh, XF).  The occupancy row written in place:
// [1] sum of the label occupancies (the blank's is one minus it), [4+k] label occupancy.
__host__ __device__ inline int ctc_em_floats(int Sp) { return 4 + Sp; }
__host__ inline CtcWs ctc_ws_layout(int T, int N, int S) {
    CtcWs w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.E = ctc_em_floats(w.Sp);
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

// grid N, block 256.  meta[n] = {T_n, L_n, invalid, number of adjacent equal labels}.
__global__ void __launch_bounds__(256) ctc_prep_kernel(PrepParams p) {
    extern __shared__ __align__(16) int s_y[];
    __shared__ int s_bad, s_rank, s_rep;
    const int n = blockIdx.x;
    long long Tn = load_idx(p.in_len, n, p.len64), Ln = load_idx(p.tgt_len, n, p.len64);
    if (threadIdx.x == 0) { s_bad = (Tn < 0 || Tn > p.T || Ln < 0 || Ln > p.S) ? 1 : 0; s_rank = 0; s_rep = 0; }
    __syncthreads();
    const int L = s_bad ? 0 : (int)Ln;
    __syncthreads();           // every thread has read the length verdict before the label check below may flip s_bad
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) {
        long long y = load_idx(p.targets, (long long)n * p.tgt_stride + k, p.tgt64);
        // labels beyond L_n are kept too: star-CTC reads targets[n, L_n] (ha/star.py:46)
        if (y < 0 || y >= p.V) { if (k < L) s_bad = 1; y = 0; }
        s_y[k] = (int)y;
    }
    __syncthreads();
    // duplicate-label chains: for position k the next position holding the same label, and whether an
    // earlier one does.  One branch-free scan of all positions per k (vector loads, no dependent exits):
    // L^2 / 4 shared-memory loads per utterance, issue-bound instead of latency-bound.
    const int Lc = p.star ? min(L + 1, p.S) : L;
    const int Lc4 = (Lc + 3) & ~3;
    for (int k = p.S + threadIdx.x; k < p.Sp; k += blockDim.x) s_y[k] = -1;   // s_y has Sp = round_up(S, 4) entries
    __syncthreads();
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) {
        const int y = s_y[k];
        int nxt = 0x7fffffff, notfirst = 0;
        if (k < Lc) {
            const int4* y4 = (const int4*)s_y;
#pragma unroll 4
            for (int j4 = 0; j4 < (Lc4 >> 2); ++j4) {
                const int4 v = y4[j4];
                const int j = 4 * j4;
                const int m0 = (v.x == y) & (j < Lc), m1 = (v.y == y) & (j + 1 < Lc);
                const int m2 = (v.z == y) & (j + 2 < Lc), m3 = (v.w == y) & (j + 3 < Lc);
                notfirst |= (m0 & (j < k)) | (m1 & (j + 1 < k)) | (m2 & (j + 2 < k)) | (m3 & (j + 3 < k));
                nxt = min(nxt, (m0 && j > k) ? j : 0x7fffffff);
                nxt = min(nxt, (m1 && j + 1 > k) ? j + 1 : 0x7fffffff);
                nxt = min(nxt, (m2 && j + 2 > k) ? j + 2 : 0x7fffffff);
                nxt = min(nxt, (m3 && j + 3 > k) ? j + 3 : 0x7fffffff);
            }
        }
        p.tgt[(size_t)n * p.Sp + k] = y | (notfirst ? kNotFirst : 0);
        p.dupnext[(size_t)n * p.Sp + k] = (nxt == 0x7fffffff) ? -1 : nxt;
        // a blank must separate equal neighbours; in CTC a label 0 can only be entered from the blank before it
        // (ha/ctc.py:140: no skip into a blank-valued state): one extra frame each
        if (k >= 1 && k < L && (s_y[k - 1] == y || (!p.star && y == 0))) atomicAdd(&s_rep, 1);
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
        p.meta[n] = make_int4(s_bad ? 0 : (int)Tn, L, s_bad, s_rep);
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
        float mx = -CUDART_INF_F;
        if (VEC4 && V <= 1024) {
            // the whole row in registers (<= 8 float4 per lane): one shared-memory pass for max and sum
            const float4* r4 = (const float4*)row;
            const int V4 = V >> 2;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = lane + 32 * i;
                v[i] = (c < V4) ? r4[c] : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
            mx = warp_max(mx);
            if (p.from_logits) {
                const float m2 = mx * kLog2e;
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s0 += ex2f(fmaf(v[i].x, kLog2e, -m2)) + ex2f(fmaf(v[i].y, kLog2e, -m2));
                    s1 += ex2f(fmaf(v[i].z, kLog2e, -m2)) + ex2f(fmaf(v[i].w, kLog2e, -m2));
                }
                l2 = m2 + log2f(warp_sum(s0 + s1));
            }
        } else {
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
            if (p.from_logits) {
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
        }
        // Emissions are stored relative to an integer per-row shift c_t = rint(log2 p of the row's likeliest
        // class), so every stored probability is <= 2^0.5 and anything within 2^-126 of the row maximum is a
        // normal fp32.  The shifts cancel in every posterior (both sweeps see the same rows); only the loss
        // needs their sum, which the trellis adds back.  The exponent is evaluated in float-float arithmetic
        // from the fp32 logit, so the stored probability is good to ~1e-7 relative whatever |log p| is.
        const float ct = round_int(fmaf(mx, kLog2e, -l2));       // emission of the row's likeliest class
        float* erow = p.em + ((size_t)n * p.T + t) * p.E;
        if (lane == 0) {
            p.lse2[(size_t)n * p.T + t] = l2;
            float K, f;
            emission_split(row[0], l2, ct, K, f);
            *(float4*)erow = make_float4(ct, emission_linear(K, f), 0.0f, 0.0f);
        }
        for (int k = lane; k < L; k += 32) {
            float K, f;
            emission_split(row[s_tgt[k]], l2, ct, K, f);
            erow[4 + k] = emission_linear(K, f);
        }
        __syncwarp();
        if (r + nstage < nrows) issue(r + nstage);
    }
}

// --------------------------------------------------------------------------------- trellis ---
struct TrellisParams {
    int T, N;
    const int4* meta; const int* order; const int* tgt; int Sp;
    float* em; int E;          // emission rows in / occupancy rows out, in place (layout: ctc_em_floats)
    float* tr; int SPX, JWp;   // stored trellis row = [JWp slot bases][2*(L+1) floats relative to them]
    float* loss; float* loss_ws;
    int nstage, G, W;          // ring stages; frames per stage; compute warps per sweep direction
    int dir_bytes;             // shared memory per direction
};

constexpr int kMaxG = 4;
constexpr int kTrellisGuard = 2048;   // (CTC) bytes in front of the first stage: label pairs that do not exist read
                                      // up to 2 KB below their stage in phase 2 instead of being predicated off
constexpr int kEmptyBarrier = 4;   // + 4 dir + stage: "stage released" hand-over from a side's compute warps (bar.arrive,
                                   // never blocking) to its producer warp (bar.sync: sleeps in hardware, no polling)
constexpr int kPhaseBarrier = 3;   // named barrier all warps of the CTA (compute + producers) meet at between the phases;
                                   // producers and compute warps arrive from different call sites, which bar.sync
                                   // with an explicit id and thread count permits (__syncthreads would not)

// per-direction shared memory: nstage x [G emission rows | G trellis rows | G occupancy rows | G x W partials]
// then mailboxes, log Z partials and the full/empty mbarriers
// (OC = floats of one occupancy row, NP = partial sums per row and warp)
__host__ __device__ inline int trellis_stage_floats(int E, int SPX, int OC, int G, int W, int NP) {
    return G * (E + SPX + OC) + round_up(G * W * NP, 4);
}
__host__ __device__ inline int trellis_dir_bytes(int E, int SPX, int OC, int nstage, int G, int W, int NP) {
    return round_up(nstage * trellis_stage_floats(E, SPX, OC, G, W, NP) * 4 + 2 * W * 16 + W * 16 + 2 * nstage * 8, 128);
}


// The producer warp of one sweep side (shared by the CTC and star-CTC trellis kernels): lane 0 issues
// every bulk copy.  Group k (counted over both phases) lives in stage k % nstage and holds up to G
// consecutive frames, which are contiguous in memory whichever way the side walks time.
//   phase 1: emission rows in.   phase 2: emission rows + the other side's stored trellis rows in,
//   occupancy rows (with the NP per-row partial sums of the W compute warps folded into floats
//   [1, 1+NP)) written back over the emission rows once the compute warps release the stage.
// PER_LANE: the compute warps leave their partial sums un-reduced (32 lanes each) and this otherwise
// idle warp does the reduction, off the compute warps' critical path.
template <int NP, bool PER_LANE>
__device__ __forceinline__ void trellis_producer(float* stages, int SF_, uint64_t* full,
                                                 int nstage, int G, int W, int E, int SPX, int OC,
                                                 float* em_base, float* tr_base, uint32_t occ_bytes,
                                                 int Tn, int steps1, int dir, int lane) {
    const int steps2 = Tn - steps1;
    const int ng1 = (steps1 + G - 1) / G, ng2 = (steps2 + G - 1) / G;
    auto group_rows = [&](int phase, int k, int& t_lo, int& cnt) {
        const int i0 = (phase ? steps1 : 0) + k * G;                       // first step of the group
        cnt = min(G, (phase ? Tn : steps1) - i0);
        t_lo = dir ? Tn - i0 - cnt : i0;
    };
    auto drain = [&](int s, int k2) {       // write back the occupancy rows of phase-2 group k2
        int t_lo, cnt;
        group_rows(1, k2, t_lo, cnt);
        float* st = stages + s * SF_;
        float* occ = st + G * (E + SPX);
        const float* ps = occ + G * OC;
        if (PER_LANE) {
            for (int r = 0; r < cnt; ++r) {
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    float b = 0.0f;
                    for (int x = 0; x < W; ++x) b += ps[((r * W + x) * NP + c) * 32 + lane];
                    b = warp_sum(b);
                    if (lane == 0) occ[r * OC + 1 + c] = b;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (!PER_LANE) {
                for (int r = 0; r < cnt; ++r) {
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        float b = 0.0f;
                        for (int x = 0; x < W; ++x) b += ps[(r * W + x) * NP + c];
                        occ[r * OC + 1 + c] = b;
                    }
                }
            }
            fence_async_smem();
            for (int r = 0; r < cnt; ++r) bulk_s2g(em_base + (size_t)(t_lo + r) * E, occ + r * OC, occ_bytes);
            bulk_commit();
            bulk_wait_read<0>();
        }
        __syncwarp();
    };
    for (int k = 0; k < ng1; ++k) {
        const int s = k % nstage, use = k / nstage;
        if (use > 0) named_bar_sync(kEmptyBarrier + 4 * dir + s, 32 * W + 32);
        if (lane == 0) {
            int t_lo, cnt;
            group_rows(0, k, t_lo, cnt);
            mbar_expect_tx(&full[s], (uint32_t)cnt * E * 4u);
            bulk_g2s(stages + s * SF_, em_base + (size_t)t_lo * E, (uint32_t)cnt * E * 4u, &full[s]);
        }
    }
    __threadfence();
    fence_async_all();
    __syncwarp();
    cta_phase_barrier(kPhaseBarrier, (int)blockDim.x);   // phase switch: both sides' stored rows are complete
    fence_async_all();
    for (int k2 = 0; k2 < ng2 + nstage; ++k2) {
        const int k = ng1 + k2;
        const int s = k % nstage, use = k / nstage;
        if (use > 0) {
            const int kprev = k - nstage;            // group that used this stage before
            if (k2 < ng2 || kprev >= ng1) named_bar_sync(kEmptyBarrier + 4 * dir + s, 32 * W + 32);
            if (kprev >= ng1) drain(s, kprev - ng1);
        }
        if (k2 < ng2 && lane == 0) {
            int t_lo, cnt;
            group_rows(1, k2, t_lo, cnt);
            float* st = stages + s * SF_;
            mbar_expect_tx(&full[s], (uint32_t)cnt * (E + SPX) * 4u);
            bulk_g2s(st, em_base + (size_t)t_lo * E, (uint32_t)cnt * E * 4u, &full[s]);
            bulk_g2s(st + G * E, tr_base + (size_t)t_lo * SPX, (uint32_t)cnt * SPX * 4u, &full[s]);
        }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// grid N (one CTA per utterance, longest first), block 32*(2W+2).  Warps [0,W) sweep alpha forward in
// time, warps [W,2W) sweep beta backward, and the two sides meet in the middle; warps 2W and 2W+1 are
// the sides' producers: one elected lane each issues every bulk (TMA) copy, so the compute warps never
// touch the copy engine.  A ring stage holds G consecutive frames: emission rows in, (phase 2) the
// other side's stored trellis rows in, occupancy rows out; a full mbarrier (TMA completion) and a named
// "released" barrier per stage hand stages over.
//
// Side d's step i is frame t = d ? T-1-i : i and it orders the label pairs its own way (beta = alpha
// on the reversed label sequence).  Warp w of a side owns J slots of 32 pairs: pair
// q = 32 (w J + j) + lane = (blank state 2q, label state 2q+1), each an extended-range linear number
// (common.cuh, XF: fp32 mantissa + int32 exponent).  A step is   [ha/ctc.py:155-167]
//     u = blank + label of pair q-1          blank' = u * p_t(blank)
//     v = label + (skip allowed ? u : blank) label' = v * p_t(label)
// The only value crossing a lane boundary is the label state of pair q-1: a shuffle inside a warp, an
// 8-byte mailbox between warps, with one named barrier per step per side.  Phase 1 (first half of the
// frames) stores every row as 32-bit words (20 mantissa bits + 12 bits of exponent distance below the
// slot's maximum, refreshed every step); phase 2 multiplies the live pre-emission sums u, v with the row
// the other side stored for the same frame: occupancy = u * beta / Z needs no division and posteriors
// take T sequential steps, not 2T.
template <int J, int W>
__global__ void __launch_bounds__(32 * (2 * W + 2)) ctc_trellis_kernel(TrellisParams p) {
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
    // An alignment exists iff there is a frame per label plus a blank between equal neighbours (every
    // stored emission is a positive number): decided here, because in the linear domain a void state next to
    // a real one picks up 2^-127 of it instead of staying exactly void.
    if (mt.z || Tn == 0 || Tn < L + mt.w) {
        const float v = mt.z ? CUDART_NAN_F : ((L == 0) ? 0.0f : CUDART_INF_F);
        if (threadIdx.x == 0) { p.loss[n] = v; p.loss_ws[n] = v; }
        return;
    }
    const int P = L + 1;
    const int E = p.E, SPX = p.SPX, JWp = p.JWp, OC = 4 + p.Sp;
    const int SF_ = trellis_stage_floats(E, SPX, OC, G, W, 32);

    unsigned char* db = smem_raw + kTrellisGuard + (size_t)dir * p.dir_bytes;
    float* stages = (float*)db;                                   // stage s: + s * SF_
    int2* mail = (int2*)(stages + nstage * SF_);                  // [2][W] label state of a warp's last pair
    double* redd = (double*)(mail + 2 * W);                       // [W] Z partial sums
    int* redi = (int*)(redd + W);                                 // [W] Z partial exponent maxima
    uint64_t* full = (uint64_t*)(redi + 2 * W);
    if (producer && lane == 0)
        for (int s = 0; s < nstage; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
    __syncthreads();

    float* em_base = p.em + (size_t)n * p.T * E;
    float* tr_base = p.tr + (size_t)n * p.T * SPX;
    const uint32_t occ_bytes = (uint32_t)(4 + round_up(L, 4)) * 4u;        // header + label occupancies
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;         // phase-1 steps of my side

    if (producer) {
        trellis_producer<1, true>(stages, SF_, full, nstage, G, W, E, SPX, OC, em_base, tr_base, occ_bytes,
                            Tn, steps1, dir, lane);
        return;
    }

    // ------------------------------------------------------------------------ compute warps ---
    const bool leader = (w == 0 && lane == 0);
    constexpr int nthr = 32 * W;                  // compute threads of my side
    // one named barrier per step per side (mailboxes change hands); immediate operands
    auto side_barrier = [&]() {
        if (dir) asm volatile("bar.sync 2, %0;" ::"n"(nthr) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"n"(nthr) : "memory");
    };
    // per slot: do my pair / my label exist, and is the skip transition into my label allowed
    // (ha/ctc.py:140-142)?
    const int q0 = 32 * (w * J) + lane;           // my pair in slot 0
    unsigned allowed = 0, hasp = 0, hasl = 0;
    {
        const int* y = p.tgt + (size_t)n * p.Sp;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int q = q0 + 32 * j;
            if (q < P) hasp |= 1u << j;
            if (q < L) hasl |= 1u << j;
            if (q >= 1 && q < L) {
                const int ycur = y[dir ? L - 1 - q : q] & kLabelMask;
                const int yprv = y[dir ? L - q : q - 1] & kLabelMask;
                const int dst = dir ? yprv : ycur;     // the label entered in forward time
                if (dst != 0 && ycur != yprv) allowed |= 1u << j;
            }
        }
    }
    // my label of slot j sits at position pos0 +- 32 j of an emission row (reversed for beta).  Slots
    // without a label read a clamped in-range emission instead (the blank when there is no label at all):
    // their states are phantoms of real magnitude that never flow back into real states (transitions only
    // go up) and are masked out of every output.
    const int pos0 = dir ? L - 1 - q0 : q0;
    const int pstep = dir ? -32 : 32;
    int poff[J];
#pragma unroll
    for (int j = 0; j < J; ++j) poff[j] = (L > 0) ? 4 + max(0, min(pos0 + pstep * j, L - 1)) : 1;
    // the other side's copy of my blank 2q is its state 2 (L - q) = R0 - 64 j, my label one below
    const int R0 = 2 * (L - q0);
    const int B0 = R0 >> 6, B1 = (R0 - 1) >> 6;   // slots of those states: exactly j lower per slot
    // my pairs are all unreachable before step `first` (pair q needs q frames to be reached) and none of
    // them can still reach the end after step `last` (one pair per remaining frame at most): outside
    // [first, last] the warp only keeps the barriers, mailboxes and stores going
    const int first = 32 * (w * J);
    const int first1 = max(first, 1);             // step 0 initialises instead
    const int last = Tn - L + (32 * (w + 1) * J - 1) + 1;

    // blank / label state of my pair in slot j, and the sums u, v they were made from (the occupancy of
    // a state is its pre-emission sum times the other side's stored value)
    float mb[J], ml[J], su[J], sv[J];
    int eb[J], el[J], eu[J], ev[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        mb[j] = ml[j] = su[j] = sv[j] = 1.0f;
        eb[j] = el[j] = eu[j] = ev[j] = kVoidE;
    }
    float csum = 0.0f, rZ = 1.0f;
    double logZ2 = 0.0;            // log2 Z of the shifted emissions
    int eZ = 0;
    bool feasible = true;

    // ring cursor: stage and the parity of its full barrier
    int s = 0; uint32_t fpar = 0;

    // one step of the recursion on the emission row `er`
    auto advance = [&](const float* er, int i) {
        const float pb = er[1];
        float pl[J];
#pragma unroll
        for (int j = 0; j < J; ++j) pl[j] = er[poff[j]];
        if (i >= first1 && i <= last) {
            // c = label state of the pair below: lane - 1; lane 0 takes lane 31 of the slot below, or the
            // mailbox the warp below filled in the previous step
            float cm[J]; int ce[J];
            {
                float rm[J]; int re[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    rm[j] = __shfl_sync(0xffffffffu, ml[j], (lane + 31) & 31);
                    re[j] = __shfl_sync(0xffffffffu, el[j], (lane + 31) & 31);
                }
                int2 in = make_int2(__float_as_int(1.0f), kVoidE);
                if (w > 0) in = mail[((i - 1) & 1) * W + w - 1];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    cm[j] = lane ? rm[j] : (j ? rm[j ? j - 1 : 0] : __int_as_float(in.x));
                    ce[j] = lane ? re[j] : (j ? re[j ? j - 1 : 0] : in.y);
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const XF b = xf_make(mb[j], eb[j]);
                const XF u = xf_add(b, xf_make(cm[j], ce[j]));
                const bool al = (allowed >> j) & 1u;
                const XF v = xf_add(xf_make(ml[j], el[j]), xf_make(al ? u.m : b.m, al ? u.e : b.e));
                su[j] = u.m; eu[j] = u.e; sv[j] = v.m; ev[j] = v.e;
                const XF nb = xf_mul_norm(u, pb), nl = xf_mul_norm(v, pl[j]);
                mb[j] = nb.m; eb[j] = nb.e; ml[j] = nl.m; el[j] = nl.e;
            }
        } else if (i == 0 && q0 == 0) {                            // ha/ctc.py:138 (first > 0 for every other warp)
            const XF one = xf_make(1.0f, 0);
            const XF b = xf_mul_norm(one, pb);
            mb[0] = b.m; eb[0] = b.e; eu[0] = 0;
            if (L > 0) { const XF c = xf_mul_norm(one, pl[0]); ml[0] = c.m; el[0] = c.e; ev[0] = 0; }
        }
        // my last pair's label state for the warp above (read in its next step)
        if (lane == 31 && w + 1 < W) mail[(i & 1) * W + w] = make_int2(__float_as_int(ml[J - 1]), el[J - 1]);
    };
    // A phase walks its frames one ring stage (<= G consecutive frames) at a time; the stage goes back to
    // the producer after its last frame.  Row pointers advance by one row per step (backwards for beta).
    const int rsgn = dir ? -1 : 1;
    const bool mailer = (lane == 31 && w + 1 < W);
    // ---------------------------------------------------------------------------- phase 1 ---
    {
        int2* prow = (int2*)(tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX + JWp) + q0;
        int* hrow = (int*)(tr_base + (size_t)(dir ? Tn - 1 : 0) * SPX) + w * J + lane;
        const long long rstep = dir ? -(long long)SPX : (long long)SPX;
        const bool hstore = lane < J && 32 * (w * J + lane) < P;
        for (int i0 = 0; i0 < steps1; i0 += G) {
            const int cnt = min(G, steps1 - i0);
            const float* er = stages + s * SF_ + (dir ? cnt - 1 : 0) * E;
            mbar_wait(&full[s], fpar);
            for (int i = i0; i < i0 + cnt; ++i, er += rsgn * E) {
                csum += er[0];
                if (i == 0 || (i >= first && i <= last)) advance(er, i);
                else if (mailer) mail[(i & 1) * W + w] = make_int2(__float_as_int(ml[J - 1]), el[J - 1]);
                // the row is stored relative to each slot's largest exponent of this very step: every word's
                // distance is >= 0 and saturates 4095 binary orders below it
                int bsel = 0;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int mx = __reduce_max_sync(0xffffffffu, max(eb[j], el[j]));
                    if ((hasp >> j) & 1u)
                        prow[32 * j] = make_int2(xf_pack(mb[j], mx - eb[j]), xf_pack(ml[j], mx - el[j]));
                    bsel = (lane == j) ? mx : bsel;
                }
                if (hstore) *hrow = bsel;
                prow = (int2*)((int*)prow + rstep);
                hrow += rstep;
                side_barrier();
            }
            named_bar_arrive(kEmptyBarrier + 4 * dir + s, nthr + 32);
            if (++s == nstage) { s = 0; fpar ^= 1u; }
        }
    }
    // my stored rows -> visible to the other side's bulk (async-proxy) loads, and vice versa
    __threadfence();
    fence_async_all();
    cta_phase_barrier(kPhaseBarrier, (int)blockDim.x);

    // ---------------------------------------------------------------------------- phase 2 ---
    // the other side's words for my pair in slot j: [ooff - 64 j] my blank, one below my label; their slot
    // bases at [B0 - j], [B1 - j].  Pairs that do not exist read harmless words inside the guard / stage
    // and are masked out of the outputs.
    const int ooff = JWp + R0;
    // mantissa product and exponent of (my pre-emission sum) x (other side's stored value)
    auto prod = [&](const int* trow, int jj, float& g0, int& x0, float& g1, int& x1) {
        const int o0 = trow[ooff - 64 * jj], o1 = trow[ooff - 64 * jj - 1];
        const int b0 = trow[B0 - jj], b1 = trow[B1 - jj];
        g0 = su[jj] * xf_unpack_m(o0);
        g1 = sv[jj] * xf_unpack_m(o1);
        x0 = eu[jj] + b0 - xf_unpack_below(o0);
        x1 = ev[jj] + b1 - xf_unpack_below(o1);
    };
    for (int i0 = steps1; i0 < Tn; i0 += G) {
        const int cnt = min(G, Tn - i0);
        const float* stg = stages + s * SF_;
        const int r0 = dir ? cnt - 1 : 0;
        const float* er = stg + r0 * E;
        const int* trow = (const int*)(stg + G * E + r0 * SPX);
        float* ob = (float*)stg + G * (E + SPX) + r0 * OC + 4 + pos0;          // my label of slot 0 in the occupancy row
        float* ps = (float*)stg + G * (E + SPX) + G * OC + (r0 * W + w) * 32 + lane;
        mbar_wait(&full[s], fpar);
        for (int i = i0; i < i0 + cnt; ++i, er += rsgn * E, trow += rsgn * SPX, ob += rsgn * OC, ps += rsgn * (W * 32)) {
            csum += er[0];
            const bool active = (i == 0) || (i >= first && i <= last);
            if (active) advance(er, i);
            else if (mailer) mail[(i & 1) * W + w] = make_int2(__float_as_int(ml[J - 1]), el[J - 1]);
            if (i == steps1) {
                // Z = sum over my side's states at the meeting frame: exponent maximum, then a scaled sum
                int pm = 4 * kVoidE;
                float zg[2 * J]; int zx[2 * J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    prod(trow, j, zg[2 * j], zx[2 * j], zg[2 * j + 1], zx[2 * j + 1]);
                    if (!(active && ((hasp >> j) & 1u))) zx[2 * j] = 4 * kVoidE;
                    if (!(active && ((hasl >> j) & 1u))) zx[2 * j + 1] = 4 * kVoidE;
                    pm = max(pm, max(zx[2 * j], zx[2 * j + 1]));
                }
                pm = __reduce_max_sync(0xffffffffu, pm);
                if (lane == 0) redi[w] = pm;
                side_barrier();
                for (int x = 0; x < W; ++x) pm = max(pm, redi[x]);
                feasible = pm > kVoidETest;
                double sm = 0.0;
#pragma unroll
                for (int j = 0; j < 2 * J; ++j) sm += (double)xf_scale(zg[j], max(zx[j] - pm, -126));
#pragma unroll
                for (int o = 16; o; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                if (lane == 0) redd[w] = sm;
                side_barrier();
                sm = 0.0;
                for (int x = 0; x < W; ++x) sm += redd[x];
                // Z = sm * 2^pm = mZ * 2^eZ with mZ in [1, 2)
                const int ex = feasible ? ilogb(sm) : 0;
                rZ = feasible ? (float)(1.0 / scalbn(sm, -ex)) : 1.0f;
                eZ = feasible ? pm + ex : (1 << 29);          // infeasible: every occupancy underflows to ~0
                logZ2 = feasible ? (double)pm + log2(sm) : 0.0;
            }
            // Only the label states' occupancies are formed: the occupancies of a frame sum to one, so the gradient
            // kernel takes the blank column as 1 - (sum of the label occupancies), which this row carries in [1].
            float bsum = 0.0f;
            if (active) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int o1 = trow[ooff - 64 * j - 1], b1 = trow[B1 - j];
                    float g1 = sv[j] * xf_unpack_m(o1);
                    g1 = xf_scale(g1 * rZ, max(ev[j] + b1 - xf_unpack_below(o1) - eZ, -126));
                    if ((hasl >> j) & 1u) { ob[pstep * j] = g1; bsum += g1; }
                }
            } else {
                // none of my states is on any complete path at this frame: their posteriors are exactly zero
#pragma unroll
                for (int j = 0; j < J; ++j)
                    if ((hasl >> j) & 1u) ob[pstep * j] = 0.0f;
            }
            *ps = bsum;         // per-lane label occupancy; the producer warp sums the 32 W of a row into float [1]
            if (i == i0 + cnt - 1) fence_async_smem();   // this group's occupancy writes -> the producer's bulk stores
            side_barrier();
        }
        named_bar_arrive(kEmptyBarrier + 4 * dir + s, nthr + 32);
        if (++s == nstage) { s = 0; fpar ^= 1u; }
    }
    if (dir == 0 && leader) {
        // log Z of the true emissions = log Z of the shifted ones + the sum of all T row shifts
        const float v = feasible ? (float)(-(logZ2 + (double)csum) * kLn2) : CUDART_INF_F;
        p.loss[n] = v; p.loss_ws[n] = v;
    }
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
    const uint32_t occ_bytes = (uint32_t)(4 + round_up(L, 4)) * 4u;

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
            for (int c = lane; c < L + 4; c += 32) dst[V + c] = osrc[c];
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
                float s = occ[4 + k];
                for (int j = s_nxt[k]; j >= 0; j = s_nxt[j]) s += occ[4 + j];
                row[w & kLabelMask] -= g * s;
            }
        }
        __syncwarp();
        if (lane == 0) row[0] -= g * (1.0f - occ[1]);      // occ[1] = sum of the label occupancies of the frame
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

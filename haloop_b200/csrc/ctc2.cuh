// ctc2.cuh — fused CTC loss + logit gradient for sm_100a, round 2.  Replaces ha/ctc.py:110-174
// (ctc_forward_score3) and its autograd backward with TWO kernels that never write emissions or
// occupancies to HBM:
//
//   ctc2_fwd_kernel  one CTA per (utterance, sweep direction).  Row warps stream the logit rows of
//                    their side's half of the frames through a TMA ring, form the row log-sum-exp
//                    and gather the blank + label probabilities straight into a shared-memory slot;
//                    the trellis warps consume the slots, advance alpha (direction 0, forward in
//                    time) or beta (direction 1, backward) and store only the LABEL states of each
//                    frame (4 floats + one exponent per group of 4 labels).  The two sides meet in
//                    the middle: whichever CTA finishes second forms Z from both boundaries.
//   ctc2_bwd_kernel  same decomposition for the other half of each side's frames: row warps reload
//                    the logit row (needed for the softmax anyway), gather the emissions from it,
//                    the trellis warps advance the recursion, multiply the live pre-emission sums
//                    with the rows the OTHER side stored in the forward pass and leave the label
//                    occupancies in shared memory; the row warp that owns the frame turns its row
//                    into (softmax - occupancy) * grad_out in place and bulk-stores it.
//
// HBM traffic per frame: logits read twice (4V + 4V), gradient written once (4V), stored label states
// written once and read once (~5 L bytes each): 12.6 V at L = 0.3 V against 22 V for the three-kernel
// path of round 1 (ctc.cuh: emission rows, packed alpha/beta rows of every state, occupancy rows).
//
// Numbers: "pair-normalised" linear domain.  A lane owns 4 consecutive (blank, label) pairs; the two states
// of a pair are plain fp32 values sharing one int32 exponent, rescaled after every step so that the larger of
// the two sits in [2^32, 2^33).  The sum-product step of a pair is 2 FADD + 1 SEL + 2 FMUL plus one exponent
// alignment of the label state below, every rounding is 6e-8 relative whatever T and log V are, and the only
// value that can be lost is a state more than 2^158 (110 nats) below the OTHER state of its own pair.  (One
// exponent per lane — 8 adjacent states — is not enough: with T >> L the alpha of adjacent states differs by
// tens of nats per label and the low states still carry posterior mass through a large beta; tools/ctc2_proto.py,
// the numpy model of this file, shows it at T = 3000, L = 4.)
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hab {

constexpr int kJP = 4;            // pairs per lane
constexpr int kR2 = 4;            // row warps per CTA
constexpr int kNE = 8;            // emission / occupancy slots per CTA (a power of two, multiple of kR2)
constexpr int kNSR = 4;           // stored-row slots (backward; a power of two)
constexpr int kRcap = 8;          // later occurrences of a class are scattered in parallel up to this occurrence rank
constexpr int kLaneExp = 32 + 127;        // biased exponent the lane maximum is normalised to
constexpr int kAlignMax = 30;             // largest up-shift applied to the neighbour's label state

struct Ctc2Ws {                   // workspace layout (byte offsets)
    size_t meta, order, tgt, nflist, nfhdr, loss, zinfo, cnt, lse2, bound, tr, total;
    int Sp, NLmax, SPL, BW, NF;
};
__host__ inline Ctc2Ws ctc2_ws_layout(int T, int N, int S) {
    Ctc2Ws w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.NLmax = w.Sp / 4 + 2;                     // lanes (groups of 4 label slots) per side
    w.SPL = 8 * w.NLmax;                        // stored row: [4 NL label values][4 NL pair exponents]
    w.BW = 12 * w.NLmax;                        // boundary: [4 NL blanks][4 NL labels][4 NL pair exponents]
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.NF = w.Sp + 32 * kRcap;                   // later occurrences of a class, padded per occurrence rank
    w.nflist = take(sizeof(int) * (size_t)N * w.NF);
    w.nfhdr = take(sizeof(int2) * (size_t)N);
    w.loss = take(sizeof(float) * (size_t)N);
    w.zinfo = take(sizeof(int4) * (size_t)N);   // {exponent of Z, bits of 1 / mantissa of Z, -, -}
    w.cnt = take(sizeof(int) * (size_t)N);      // arrivals at the meeting point
    w.lse2 = take(sizeof(float) * (size_t)N * T);
    w.bound = take(sizeof(int) * (size_t)N * 2 * w.BW);
    w.tr = take(sizeof(int) * (size_t)N * T * w.SPL);
    w.total = o;
    return w;
}

struct Ctc2Params {
    const float* x; long long sx_t, sx_n;
    float* gx; long long sg_t, sg_n;
    int T, N, V, Sp;
    const int4* meta; const int* order; const int* tgt; const int* nflist; const int2* nfhdr; int NF;
    float* lse2; int* tr; int SPL; int* bound; int BW; int4* zinfo; int* cnt;
    float* loss; float* loss_ws; const float* gout;
    int from_logits, NS, NLmax, EMF;      // ring stages per row warp (1 or 2); lanes per side of the longest target;
                                          // floats per emission slot (ctc2_em_floats)
};

// An emission slot: [0] blank, [4 + a] label index a = position + 4 (a < 4 and a >= L + 4 are phantoms that stay 0).
// Sized so that the lanes of the longest target and a row warp's four-labels-per-lane batches stay inside it.
__host__ __device__ inline int ctc2_em_floats(int Sp) { return max(Sp + 12, 8 + round_up(Sp, 128)); }

// shared memory (bytes): [mbarriers][mailboxes][Z reduction][targets (+ later occurrences)][emission slots][stored slots][row ring]
struct Ctc2Smem { int bars, mail, red, tgt, em, st, rows, total; };
__host__ __device__ inline Ctc2Smem ctc2_smem(int W, int NS, int V, int Sp, int SPL, bool bwd) {
    Ctc2Smem s;
    int o = 0;
    auto take = [&](int bytes) { int at = o; o = round_up(o + bytes, 128); return at; };
    s.bars = take(8 * (kR2 * NS + 2 * kNE + kNSR));
    s.mail = take(8 * 2 * W);
    s.red = take(16 * W + 16);
    s.tgt = take(4 * round_up(Sp, 128));
    s.em = take(4 * kNE * ctc2_em_floats(Sp));
    s.st = take(bwd ? 4 * kNSR * SPL : 0);
    s.rows = take(4 * kR2 * NS * V);
    s.total = o;
    return s;
}

// ---------------------------------------------------------------------------------- emissions ---
// p = 2^(x log2e - l2) for a logit x of a row whose log2-sum-exp2 is l2.  l2 is split into a multiple of
// 2^-10 (l2q) and a remainder (dl) so that the integer part of the exponent can be removed BEFORE the one
// rounding that matters: the fraction handed to ex2 is exact to 3e-8 whatever |log p| is.  The forward and
// the backward kernel call this with bit-identical arguments, so both halves of a trellis see one model.
// Floor: 2^-125.75 (a class more than 87 nats below probability one), so every emission is a positive normal.
struct RowNorm { float l2q, dl; };
__device__ __forceinline__ RowNorm row_norm(float l2) {
    RowNorm r;
    r.l2q = rintf(l2 * 1024.0f) * (1.0f / 1024.0f);
    r.dl = l2 - r.l2q;
    return r;
}
__device__ __forceinline__ float emission2(float x, RowNorm rn) {
    const float t = fmaxf(fmaf(x, kLog2e, -rn.l2q), -125.0f);
    const float tk = t + kMagic;                                   // integer part k in the low mantissa bits
    const float kf = tk - kMagic;
    const float f = fmaxf(fmaf(x, kLog2e, -(rn.l2q + kf)) - rn.dl, -0.75f);    // |f| <= 0.5 + |dl|; the clamp only acts below the floor
    // p * 2^k by exponent-field arithmetic: bits(tk) = 0x4B400000 + k and 0x4B400000 << 23 == 0 (mod 2^32)
    return __int_as_float(__float_as_int(ex2f(f)) + (__float_as_int(tk) << 23));
}

// Blank + label emissions of one row into an emission slot: four independent labels per lane and iteration, so
// the LDS -> LDS -> FFMA -> MUFU -> STS chains of a row overlap.  s_off: byte offset of the label's logit in the
// row (label * 4, flag bits below), padded with zeros to a multiple of 128 entries.
__device__ __forceinline__ void gather_row(float* em, const float* row, const int* s_off, int L, int lane, RowNorm rn) {
    if (lane == 0) em[0] = emission2(row[0], rn);
    for (int k0 = lane; k0 < L; k0 += 128) {
        float xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = *(const float*)((const char*)row + (s_off[k0 + 32 * q] & ~3));
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q] = emission2(xv[q], rn);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k0 + 32 * q < L) em[8 + k0 + 32 * q] = xv[q];
    }
}

// log2-sum-exp2 of a row of V logits (V % 4 == 0) held in shared memory.  One pass without the usual maximum
// when the plain sum stays inside the fp32 range (|logit| < ~85); otherwise the shifted two-pass form.
__device__ __forceinline__ float row_lse2(const float* row, int V4, int lane) {
    const float4* r4 = (const float4*)row;
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 4
    for (int c = lane; c < V4; c += 32) {
        const float4 v = r4[c];
        s0 += ex2f(v.x * kLog2e) + ex2f(v.y * kLog2e);
        s1 += ex2f(v.z * kLog2e) + ex2f(v.w * kLog2e);
    }
    const float s = warp_sum(s0 + s1);
    if (s > 1e-30f && s < 1e30f) return log2f(s);
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
        s0 += ex2f(fmaf(v.x, kLog2e, -m2)) + ex2f(fmaf(v.y, kLog2e, -m2));
        s1 += ex2f(fmaf(v.z, kLog2e, -m2)) + ex2f(fmaf(v.w, kLog2e, -m2));
    }
    return m2 + log2f(warp_sum(s0 + s1));
}

// -------------------------------------------------------------------------------- lane numbers ---
// A lane's 4 pairs are kept in POSITION order for both directions: component c is the pair whose label index is
// a = 4 g + c.  Direction 1 (beta) walks them downwards, so "the pair below" is c + 1 and the value entering
// from the previous lane of the side lands at c = 3; nothing is reversed when rows are loaded or stored.
struct Lane { float b[kJP], l[kJP]; int e[kJP]; };

// x * 2^d for x >= 0 by exponent-field arithmetic: exact while the result is a normal number, 0 below 2^-126
// (a factor 2^d formed on its own would already be 0 at d < -126, although x * 2^d is not).  d <= 64.
__device__ __forceinline__ float scale_pow2(float x, int d) {
    const int b = __float_as_int(x);
    return ((b >> 23) + d > 0) ? __int_as_float(b + (d << 23)) : 0.0f;
}

// First half of a step: align the label state of the pair below each of my pairs (cm, ce: the one entering from
// the previous lane) to the pair's exponent and form the pre-emission sums
//     u = blank + label below,  v = label + (skip allowed ? u : blank)                 [ha/ctc.py:155-167]
// Their scale is s.e[c] on return.
template <int DIR>
__device__ __forceinline__ void lane_sums(Lane& s, unsigned allowed, float cm, int ce, float (&u)[kJP], float (&v)[kJP]) {
    float lb[kJP]; int d[kJP];
#pragma unroll
    for (int c = 0; c < kJP; ++c) {
        const bool edge = DIR ? (c == kJP - 1) : (c == 0);
        const int cb = DIR ? (c < kJP - 1 ? c + 1 : c) : (c ? c - 1 : 0);
        lb[c] = edge ? cm : s.l[cb];
        d[c] = (edge ? ce : s.e[cb]) - s.e[c];
    }
    if (max(max(d[0], d[1]), max(d[2], d[3])) > kAlignMax) {        // a pair dwarfed by the one below it (the wavefront
#pragma unroll                                                      // arrives): move its exponent up
        for (int c = 0; c < kJP; ++c) {
            if (d[c] > kAlignMax) {
                const int sh = d[c] - kAlignMax;
                s.b[c] = scale_pow2(s.b[c], -min(sh, 512)); s.l[c] = scale_pow2(s.l[c], -min(sh, 512)); s.e[c] += sh;
                d[c] = kAlignMax;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kJP; ++c) {
        u[c] = s.b[c] + scale_pow2(lb[c], max(d[c], -512));
        v[c] = s.l[c] + (((allowed >> c) & 1u) ? u[c] : s.b[c]);
    }
}
// Second half: multiply by the emissions and renormalise each pair.
__device__ __forceinline__ void lane_emit(Lane& s, const float (&u)[kJP], const float (&v)[kJP], float pb, float4 pl) {
    const float plc[kJP] = {pl.x, pl.y, pl.z, pl.w};
#pragma unroll
    for (int c = 0; c < kJP; ++c) {
        const float nb = u[c] * pb, nl = v[c] * plc[c];
        const float mx = fmaxf(nb, nl);
        const int delta = min(kLaneExp - (__float_as_int(mx) >> 23), 120);
        const float f = __int_as_float((delta + 127) << 23);
        s.b[c] = nb * f; s.l[c] = nl * f;
        s.e[c] = (mx > 0.0f) ? s.e[c] - delta : kVoidE;
    }
}

// Per-lane constants of a trellis thread.
struct LaneCfg {
    int gl;            // lane index within the side (32 w + lane); direction 1 counts the groups downwards
    int g;             // my position group (clamped into the slot): labels a = 4 g .. 4 g + 3
    unsigned allowed;  // skip-transition bits of my 4 pairs, by component   [ha/ctc.py:140-142]
    bool live;         // gl < NL: the lane owns a group of the stored rows
};
__device__ __forceinline__ LaneCfg lane_cfg(int gl, int dir, int L, int NL, int NLmax, const int* y) {
    LaneCfg cf;
    cf.gl = gl;
    cf.live = gl < NL;
    const int g = dir ? NL - 1 - gl : gl;
    cf.g = min(max(g, 0), NLmax - 1);
    cf.allowed = 0;
#pragma unroll
    for (int c = 0; c < kJP; ++c) {
        const int p = 4 * g + c - 4;                    // label position
        bool al = false;
        if (!dir) {
            if (p == 0) al = true;
            else if (p >= 1 && p < L) { const int yc = y[p] & kLabelMask, yp = y[p - 1] & kLabelMask; al = (yc != yp) && (yc != 0); }
        } else {
            if (p == L - 1 || p == -1) al = true;       // p == -1: the last pair's phantom label only ever collects Z
            else if (p >= 0 && p + 1 < L) { const int yc = y[p] & kLabelMask, yn = y[p + 1] & kLabelMask; al = (yn != yc) && (yn != 0); }
        }
        if (al) cf.allowed |= 1u << c;
    }
    return cf;
}

__device__ __forceinline__ void side_barrier(int nthr) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
}

// the label state of the pair below my lowest one: the previous lane of the side, or the mailbox the warp below
// filled last step
template <int DIR>
__device__ __forceinline__ void fetch_below(const Lane& s, int w, int lane, const int2* mail_prev, float& cm, int& ce) {
    cm = __shfl_up_sync(0xffffffffu, s.l[DIR ? 0 : kJP - 1], 1);
    ce = __shfl_up_sync(0xffffffffu, s.e[DIR ? 0 : kJP - 1], 1);
    if (lane == 0) {
        int2 in = make_int2(0, kVoidE);
        if (w > 0) in = mail_prev[w - 1];
        cm = __int_as_float(in.x); ce = in.y;
    }
}

// waits of warps that poll often: a short sleep between polls leaves the issue slots to the warps that have work
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        __nanosleep(ns);
        if (mbar_try_wait(bar, parity)) return;
    }
    __trap();
}

// --------------------------------------------------------------------------------------- prep ---
struct Prep2Params {
    const void* targets; long long tgt_stride; int tgt64;
    const void* in_len; const void* tgt_len; int len64;
    int T, N, V, S, Sp, NF;
    int4* meta; int* order; int* tgt; int* nflist; int2* nfhdr; int* cnt; int4* zinfo;
    int star;   // star-CTC (star2.cuh): occurrence ranks also cover position L_n (< S), whose star reads targets[n, L_n]
                // (ha/star.py:46), and a label 0 needs no extra frame
};

// grid N, block 256.  meta[n] = {T_n, L_n, invalid, frames an alignment needs beyond L_n}; tgt[n][k] = label |
// kNotFirst; nflist[n] = the positions whose label occurred before, as (label << 10 | position): occurrence ranks
// 1 .. kRcap-1 in groups padded to 32 entries (-1) — the classes of a group are distinct, so the gradient kernel
// updates one group per instruction — then any higher ranks in position order (a serial tail).
static __global__ void __launch_bounds__(256) ctc2_prep_kernel(Prep2Params p) {
    extern __shared__ __align__(16) int s_y[];          // [Sp] labels, [Sp] occurrence ranks
    __shared__ int s_bad, s_rank, s_rep, s_cnt[kRcap], s_off[kRcap], s_fill[kRcap];
    int* s_rk = s_y + p.Sp;
    const int n = blockIdx.x;
    const long long Tn = load_idx(p.in_len, n, p.len64), Ln = load_idx(p.tgt_len, n, p.len64);
    const bool lenbad = (Tn < 0 || Tn > p.T || Ln < 0 || Ln > p.S);
    if (threadIdx.x == 0) { s_bad = lenbad ? 1 : 0; s_rank = 0; s_rep = 0; }
    if (threadIdx.x < kRcap) { s_cnt[threadIdx.x] = 0; s_fill[threadIdx.x] = 0; }
    __syncthreads();
    const int L = lenbad ? 0 : (int)Ln;
    for (int k = threadIdx.x; k < p.S; k += blockDim.x) {
        long long y = load_idx(p.targets, (long long)n * p.tgt_stride + k, p.tgt64);
        if (y < 0 || y >= p.V) { if (k < L) s_bad = 1; y = 0; }
        s_y[k] = (int)y;
    }
    for (int k = p.S + threadIdx.x; k < p.Sp; k += blockDim.x) s_y[k] = -1;
    int* nfl = p.nflist + (size_t)n * p.NF;
    for (int k = threadIdx.x; k < p.NF; k += blockDim.x) nfl[k] = -1;
    __syncthreads();
    const int L4 = (L + 3) & ~3;
    const int Lc = p.star ? min(L + 1, p.S) : L;
    for (int k = threadIdx.x; k < p.Sp; k += blockDim.x) {
        const int y = s_y[k];
        int rk = 0;                                     // earlier positions with my label (branch-free vector scan)
        if (k < Lc) {
            const int4* y4 = (const int4*)s_y;
#pragma unroll 4
            for (int j4 = 0; j4 < (L4 >> 2); ++j4) {
                const int4 v = y4[j4];
                const int j = 4 * j4;
                rk += ((v.x == y) & (j < k)) + ((v.y == y) & (j + 1 < k)) + ((v.z == y) & (j + 2 < k)) + ((v.w == y) & (j + 3 < k));
            }
            if (rk > 0) atomicAdd(&s_cnt[min(rk, kRcap) - 1], 1);      // s_cnt[kRcap - 1]: the serial tail
        }
        s_rk[k] = rk;
        p.tgt[(size_t)n * p.Sp + k] = (y < 0 ? 0 : y) | (rk ? kNotFirst : 0);
        // a blank must separate equal neighbours, and a label 0 can only be entered from the blank before
        // it (ha/ctc.py:140: no skip into a blank-valued state): one extra frame each
        if (k >= 1 && k < L && (s_y[k - 1] == y || (!p.star && y == 0))) atomicAdd(&s_rep, 1);
    }
    {   // longest-first work order (rank by counting; N is a batch size)
        const long long mine = lenbad ? 0 : Tn * (Ln + 1);
        int rank = 0;
        for (int m = threadIdx.x; m < p.N; m += blockDim.x) {
            const long long Tm = load_idx(p.in_len, m, p.len64), Lm = load_idx(p.tgt_len, m, p.len64);
            const long long c = (Tm < 0 || Tm > p.T || Lm < 0 || Lm > p.S) ? 0 : Tm * (Lm + 1);
            rank += (c > mine) || (c == mine && m < n);
        }
        if (rank) atomicAdd(&s_rank, rank);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int o = 0;
        for (int r = 0; r < kRcap - 1; ++r) { s_off[r] = o; o += (s_cnt[r] + 31) & ~31; }
        s_off[kRcap - 1] = o;
        p.meta[n] = make_int4(s_bad ? 0 : (int)Tn, L, s_bad, s_rep);
        p.nfhdr[n] = make_int2(o, s_cnt[kRcap - 1]);
        p.order[s_rank] = n;
        p.cnt[n] = 0;
        p.zinfo[n] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Lc; k += blockDim.x) {
        const int rk = s_rk[k];
        if (rk >= 1 && rk < kRcap) nfl[s_off[rk - 1] + atomicAdd(&s_fill[rk - 1], 1)] = (s_y[k] << 10) | k;
    }
    if (threadIdx.x == 0 && s_cnt[kRcap - 1] > 0) {     // rare: a label occurring more than kRcap times
        int o = s_off[kRcap - 1];
        for (int k = 0; k < Lc; ++k)
            if (s_rk[k] >= kRcap) nfl[o++] = (s_y[k] << 10) | k;
    }
}

// ------------------------------------------------------------------------------------ forward ---
// grid 2N (CTA c: utterance order[c / 2], direction c % 2), block 32 (W + kR2).
template <int W, int MINB>
__global__ void __launch_bounds__(32 * (W + kR2), MINB) ctc2_fwd_kernel(Ctc2Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.order[blockIdx.x >> 1], dir = blockIdx.x & 1;
    const int4 mt = p.meta[n];
    const int Tn = mt.x, L = mt.y;
    if (mt.z || Tn == 0 || Tn < L + mt.w) {       // invalid / empty / no alignment exists: decided combinatorially
        if (dir == 0 && threadIdx.x == 0) {
            const float v = mt.z ? CUDART_NAN_F : ((L == 0) ? 0.0f : CUDART_INF_F);
            p.loss[n] = v; p.loss_ws[n] = v;
        }
        return;
    }
    const int NL = (L + 3) / 4 + 2;
    const int Wn = (NL + 31) >> 5;                 // trellis warps this utterance needs
    const int NS = p.NS, nsm = NS - 1, V = p.V, EMF = p.EMF;
    const Ctc2Smem sm = ctc2_smem(W, NS, V, p.Sp, p.SPL, false);
    uint64_t* row_full = (uint64_t*)(smem + sm.bars);            // [kR2][NS]
    uint64_t* em_full = row_full + kR2 * NS;                     // [kNE]
    uint64_t* em_empty = em_full + kNE;                          // [kNE]
    int2* mail = (int2*)(smem + sm.mail);                        // [2][W]
    double* redd = (double*)(smem + sm.red);                     // [W]
    int* redi = (int*)(redd + W);                                // [W] + flag
    int* s_off = (int*)(smem + sm.tgt);
    float* s_em = (float*)(smem + sm.em);
    float* s_rows = (float*)(smem + sm.rows);
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kR2 * NS; ++i) mbar_init(&row_full[i], 1);
        for (int i = 0; i < kNE; ++i) { mbar_init(&em_full[i], 32); mbar_init(&em_empty[i], 32 * Wn); }
    }
    for (int i = threadIdx.x; i < kNE * EMF; i += blockDim.x) s_em[i] = 0.0f;
    for (int k = threadIdx.x; k < round_up(L, 128); k += blockDim.x)
        s_off[k] = (k < L) ? (p.tgt[(size_t)n * p.Sp + k] & kLabelMask) << 2 : 0;
    mbar_init_fence();
    __syncthreads();

    if (warp >= W) {
        // ---------------------------------------------------------------------- row warps ---
        const int r = warp - W;
        float* wrows = s_rows + (size_t)r * NS * V;
        uint64_t* wbar = row_full + r * NS;
        const float* xb = p.x + (long long)n * p.sx_n;
        const int nrows = (steps1 > r) ? (steps1 - 1 - r) / kR2 + 1 : 0;
        const int V4 = V >> 2;
        auto issue = [&](int k) {
            const int i = r + k * kR2, t = dir ? Tn - 1 - i : i;
            if (lane == 0) {
                mbar_expect_tx(&wbar[k & nsm], (uint32_t)V * 4u);
                bulk_g2s(wrows + (k & nsm) * V, xb + (long long)t * p.sx_t, (uint32_t)V * 4u, &wbar[k & nsm]);
            }
        };
        for (int k = 0; k < min(NS, nrows); ++k) issue(k);
        for (int k = 0; k < nrows; ++k) {
            const int i = r + k * kR2, t = dir ? Tn - 1 - i : i;
            const float* row = wrows + (k & nsm) * V;
            mbar_wait(&wbar[k & nsm], (uint32_t)(k >> nsm) & 1u);
            const float l2 = p.from_logits ? row_lse2(row, V4, lane) : 0.0f;
            if (lane == 0) p.lse2[(size_t)n * p.T + t] = l2;
            const int slot = i & (kNE - 1), use = i / kNE;
            if (use > 0) mbar_wait_sleep(&em_empty[slot], (uint32_t)(use - 1) & 1u, 128);
            gather_row(s_em + slot * EMF, row, s_off, L, lane, row_norm(l2));
            mbar_arrive(&em_full[slot]);
            __syncwarp();
            if (k + NS < nrows) issue(k + NS);
        }
        return;
    }
    if (warp >= Wn) return;

    // ------------------------------------------------------------------------ trellis warps ---
    const int w = warp;
    const int nthr = 32 * Wn;
    const LaneCfg cfg = lane_cfg(32 * w + lane, dir, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
    Lane s;
#pragma unroll
    for (int c = 0; c < kJP; ++c) { s.b[c] = 0.0f; s.l[c] = 0.0f; s.e[c] = kVoidE; }
    const int k_inj = dir ? 4 * NL - 1 - (L + 4) : 3;       // the virtual source: mass 1 on the label below the first real pair
    if (cfg.gl == (k_inj >> 2)) {
        const int cj = dir ? 3 - (k_inj & 3) : (k_inj & 3);
#pragma unroll
        for (int c = 0; c < kJP; ++c)
            if (c == cj) { s.l[c] = 1.0f; s.e[c] = 0; }
    }
    if (lane == 31) mail[1 * W + w] = make_int2(__float_as_int(s.l[dir ? 0 : kJP - 1]), s.e[dir ? 0 : kJP - 1]);
    side_barrier(nthr);
    // my pairs are all unreachable before step `first` and none can still complete after step `last`
    const int q_lo = 128 * w - k_inj - 1, q_hi = 128 * w + 127 - k_inj - 1;     // real pair indices of this warp
    const int first = max(q_lo, 0), last = Tn - L + q_hi;
    const float* emp = s_em + 4 + 4 * cfg.g;                // my four label emissions in slot 0

    auto sweep = [&](auto dirc) {
        constexpr int DIR = decltype(dirc)::value;
        int* trow = p.tr + ((size_t)n * p.T + (DIR ? Tn - 1 : 0)) * p.SPL + 4 * cfg.g;
        const int eoff = 4 * NL;                             // from my label group to my exponent group
        const long long tstep = DIR ? -(long long)p.SPL : (long long)p.SPL;
        bool dead = false;
        for (int i = 0; i < steps1; ++i) {
            const int slot = i & (kNE - 1);
            mbar_wait_sleep(&em_full[slot], (uint32_t)(i / kNE) & 1u, 32);
            const float pb = s_em[slot * EMF];
            const float4 pl = *(const float4*)(emp + slot * EMF);
            mbar_arrive(&em_empty[slot]);
            if (i >= first && i <= last) {
                float cm; int ce;
                fetch_below<DIR>(s, w, lane, mail + ((i + 1) & 1) * W, cm, ce);
                float u[kJP], v[kJP];
                lane_sums<DIR>(s, cfg.allowed, cm, ce, u, v);
                lane_emit(s, u, v, pb, pl);
            } else if (i > last && !dead) {
                dead = true;
#pragma unroll
                for (int c = 0; c < kJP; ++c) { s.b[c] = 0.0f; s.l[c] = 0.0f; s.e[c] = kVoidE; }
            }
            if (lane == 31) mail[(i & 1) * W + w] = make_int2(__float_as_int(s.l[DIR ? 0 : kJP - 1]), s.e[DIR ? 0 : kJP - 1]);
            if (cfg.live) {
                *(float4*)trow = make_float4(s.l[0], s.l[1], s.l[2], s.l[3]);
                *(int4*)(trow + eoff) = make_int4(s.e[0], s.e[1], s.e[2], s.e[3]);
            }
            trow += tstep;
            side_barrier(nthr);
        }
    };
    if (dir) sweep(std::integral_constant<int, 1>{}); else sweep(std::integral_constant<int, 0>{});

    // ---- the meeting: leave my boundary state; whoever arrives second forms Z from both ----
    int* mybound = p.bound + ((size_t)n * 2 + dir) * p.BW;
    if (cfg.live) {
        *(float4*)(mybound + 4 * cfg.g) = make_float4(s.b[0], s.b[1], s.b[2], s.b[3]);
        *(float4*)(mybound + 4 * NL + 4 * cfg.g) = make_float4(s.l[0], s.l[1], s.l[2], s.l[3]);
        *(int4*)(mybound + 8 * NL + 4 * cfg.g) = make_int4(s.e[0], s.e[1], s.e[2], s.e[3]);
    }
    __threadfence();
    side_barrier(nthr);
    if (threadIdx.x == 0) { redi[W] = atomicAdd(&p.cnt[n], 1); __threadfence(); }
    side_barrier(nthr);
    if (redi[W] == 0) return;                                  // the other side is still sweeping: it will do it
    {
        // Z = sum over the states of (alpha's pre-emission sums of the meeting frame) x (beta's boundary).  Whichever
        // side gets here runs the SAME arithmetic on the two stored boundaries, so the result does not depend on
        // which CTA finished last (bit-identical repeats).
        const int* ba = p.bound + ((size_t)n * 2 + 0) * p.BW;
        const int* ob = p.bound + ((size_t)n * 2 + 1) * p.BW;
        const LaneCfg ca = lane_cfg(32 * w + lane, 0, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
        Lane sa;
#pragma unroll
        for (int c = 0; c < kJP; ++c) {
            sa.b[c] = ca.live ? __int_as_float(__ldcg(ba + 4 * ca.g + c)) : 0.0f;
            sa.l[c] = ca.live ? __int_as_float(__ldcg(ba + 4 * NL + 4 * ca.g + c)) : 0.0f;
            sa.e[c] = ca.live ? __ldcg(ba + 8 * NL + 4 * ca.g + c) : kVoidE;
        }
        if (lane == 31) mail[w] = make_int2(__float_as_int(sa.l[kJP - 1]), sa.e[kJP - 1]);
        side_barrier(nthr);
        float cm; int ce;
        float u[kJP], v[kJP];
        fetch_below<0>(sa, w, lane, mail, cm, ce);
        lane_sums<0>(sa, ca.allowed, cm, ce, u, v);
        // alpha's blank of the pair with label index a is beta's blank stored with label a - 1
        const int A = 4 * NL;
        float zm[2 * kJP]; int zx[2 * kJP];
        int pm = 4 * kVoidE;
#pragma unroll
        for (int c = 0; c < kJP; ++c) {
            const int a = 4 * ca.g + c, ab = a - 1;
            float mb = 0.0f, ml = 0.0f; int xb = 0, xl = 0;
            if (ca.live) {
                ml = v[c] * __int_as_float(__ldcg(ob + 4 * NL + a));
                xl = sa.e[c] + __ldcg(ob + 8 * NL + a);
                if (ab >= 0 && ab < A) {
                    mb = u[c] * __int_as_float(__ldcg(ob + ab));
                    xb = sa.e[c] + __ldcg(ob + 8 * NL + ab);
                }
            }
            zm[2 * c] = mb; zx[2 * c] = (mb > 0.0f) ? xb + (__float_as_int(mb) >> 23) : 4 * kVoidE;
            zm[2 * c + 1] = ml; zx[2 * c + 1] = (ml > 0.0f) ? xl + (__float_as_int(ml) >> 23) : 4 * kVoidE;
            pm = max(pm, max(zx[2 * c], zx[2 * c + 1]));
        }
        // pm: the largest (scale + biased fp32 exponent) of any term; terms are summed relative to it in float64
        pm = __reduce_max_sync(0xffffffffu, pm);
        if (lane == 0) redi[w] = pm;
        side_barrier(nthr);
        for (int x = 0; x < Wn; ++x) pm = max(pm, redi[x]);
        const bool feasible = pm > kVoidETest;
        double sum = 0.0;
#pragma unroll
        for (int j = 0; j < 2 * kJP; ++j) {
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
// grid 2N, block 32 (W + kR2).  d loss / d logits = (softmax - occupancy) * gout (from_logits) or
// -occupancy * gout (log-prob input, the reference's autograd boundary); rows t >= T_n and every row of an
// infeasible or invalid utterance are zero.
template <int W, int MINB>
__global__ void __launch_bounds__(32 * (W + kR2), MINB) ctc2_bwd_kernel(Ctc2Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.order[blockIdx.x >> 1], dir = blockIdx.x & 1;
    const int4 mt = p.meta[n];
    const int L = mt.y;
    const float lossn = p.loss_ws[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x;      // NaN / inf loss: all-zero gradient
    const int NS = p.NS, nsm = NS - 1, V = p.V, EMF = p.EMF, V4 = V >> 2;
    float* gb = p.gx + (long long)n * p.sg_n;
    const int tm = Tn >> 1;
    const int steps1 = dir ? Tn - tm : tm;
    const int nsteps2 = Tn - steps1;

    if (warp >= W) {
        // rows past the end of the utterance: zero, shared between the two sides' row warps
        const int r = warp - W;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = Tn + dir + 2 * r; t < p.T; t += 2 * kR2) {
            float4* dst = (float4*)(gb + (long long)t * p.sg_t);
            for (int c = lane; c < V4; c += 32) dst[c] = z;
        }
    }
    if (nsteps2 <= 0) return;

    const int NL = (L + 3) / 4 + 2;
    const int Wn = (NL + 31) >> 5;
    const Ctc2Smem sm = ctc2_smem(W, NS, V, p.Sp, p.SPL, true);
    uint64_t* row_full = (uint64_t*)(smem + sm.bars);            // [kR2][NS]
    uint64_t* em_full = row_full + kR2 * NS;                     // [kNE]
    uint64_t* occ_full = em_full + kNE;                          // [kNE]
    uint64_t* st_full = occ_full + kNE;                          // [kNSR]
    int2* mail = (int2*)(smem + sm.mail);
    int* s_off = (int*)(smem + sm.tgt);                          // label * 4 | "occurred before", padded to 128 entries
    const int* nfl = p.nflist + (size_t)n * p.NF;                // later occurrences of a class (ctc2_prep_kernel)
    float* s_em = (float*)(smem + sm.em);
    int* s_st = (int*)(smem + sm.st);
    float* s_rows = (float*)(smem + sm.rows);
    const float g = p.gout[n];

    if (threadIdx.x == 0) {
        for (int i = 0; i < kR2 * NS; ++i) mbar_init(&row_full[i], 1);
        for (int i = 0; i < kNE; ++i) { mbar_init(&em_full[i], 32); mbar_init(&occ_full[i], 32 * Wn); }
        for (int i = 0; i < kNSR; ++i) mbar_init(&st_full[i], 1);
    }
    for (int i = threadIdx.x; i < kNE * EMF; i += blockDim.x) s_em[i] = 0.0f;
    const int2 nf = p.nfhdr[n];                                  // {entries in rank groups of 32, serial tail}
    for (int k = threadIdx.x; k < round_up(L, 128); k += blockDim.x) {
        const int wd = (k < L) ? p.tgt[(size_t)n * p.Sp + k] : kNotFirst;
        s_off[k] = ((wd & kLabelMask) << 2) | ((wd & kNotFirst) ? 1 : 0);
    }
    mbar_init_fence();
    __syncthreads();

    if (warp >= W) {
        // ---------------------------------------------------------------------- row warps ---
        const int r = warp - W;
        float* wrows = s_rows + (size_t)r * NS * V;
        uint64_t* wbar = row_full + r * NS;
        const float* xb = p.x + (long long)n * p.sx_n;
        const int nrows = (nsteps2 > r) ? (nsteps2 - 1 - r) / kR2 + 1 : 0;
        auto frame = [&](int k) { const int i = steps1 + r + k * kR2; return dir ? Tn - 1 - i : i; };
        auto issue = [&](int k) {
            if (lane == 0) {
                mbar_expect_tx(&wbar[k & nsm], (uint32_t)V * 4u);
                bulk_g2s(wrows + (k & nsm) * V, xb + (long long)frame(k) * p.sx_t, (uint32_t)V * 4u, &wbar[k & nsm]);
            }
        };
        // pre(k): gather the emissions of row k for the trellis warps, then turn the row into softmax * g in place.
        auto pre = [&](int k) {
            const int i2 = r + k * kR2, t = frame(k);
            float* row = wrows + (k & nsm) * V;
            const float l2 = p.lse2[(size_t)n * p.T + t];
            mbar_wait(&wbar[k & nsm], (uint32_t)(k >> nsm) & 1u);
            // (slot i2 % kNE was last used by my own row k - kNE / kR2, whose occupancies I consumed in program order)
            gather_row(s_em + (i2 & (kNE - 1)) * EMF, row, s_off, L, lane, row_norm(l2));
            mbar_arrive(&em_full[i2 & (kNE - 1)]);
            __syncwarp();
            float4* r4 = (float4*)row;
            if (p.from_logits) {
#pragma unroll 4
                for (int c = lane; c < V4; c += 32) {
                    float4 v = r4[c];
                    v.x = g * ex2f(fmaf(v.x, kLog2e, -l2)); v.y = g * ex2f(fmaf(v.y, kLog2e, -l2));
                    v.z = g * ex2f(fmaf(v.z, kLog2e, -l2)); v.w = g * ex2f(fmaf(v.w, kLog2e, -l2));
                    r4[c] = v;
                }
            } else {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = lane; c < V4; c += 32) r4[c] = z;
            }
            __syncwarp();
        };
        // post(k): subtract the occupancies the trellis warps left in the slot and store the row.
        auto post = [&](int k) {
            const int i2 = r + k * kR2, t = frame(k);
            float* row = wrows + (k & nsm) * V;
            const float* em = s_em + (i2 & (kNE - 1)) * EMF;
            mbar_wait_sleep(&occ_full[i2 & (kNE - 1)], (uint32_t)(i2 / kNE) & 1u, 128);
            // first occurrences of a class: all distinct, four per lane and iteration (the padding is "not first" with
            // occupancy zero, so no bounds are checked)
            float bs = 0.0f;
            for (int k0 = lane; k0 < L; k0 += 128) {
                int wd[4]; float oc[4], rv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { wd[q] = s_off[k0 + 32 * q]; oc[q] = em[8 + k0 + 32 * q]; }
#pragma unroll
                for (int q = 0; q < 4; ++q) { rv[q] = *(const float*)((const char*)row + (wd[q] & ~3)); bs += oc[q]; }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (!(wd[q] & 1)) *(float*)((char*)row + wd[q]) = fmaf(-g, oc[q], rv[q]);
            }
            __syncwarp();
            // later occurrences: the list holds them by occurrence rank, one rank (distinct classes) per 32 entries,
            // so a class is updated in position order by one lane at a time: deterministic, no atomics
            for (int e0 = 0; e0 < nf.x; e0 += 32) {
                const int e = __ldg(nfl + e0 + lane);
                if (e >= 0) row[e >> 10] = fmaf(-g, em[8 + (e & 1023)], row[e >> 10]);
                __syncwarp();
            }
            if (nf.y > 0 && lane == 0)
                for (int x = nf.x; x < nf.x + nf.y; ++x) { const int e = __ldg(nfl + x); row[e >> 10] = fmaf(-g, em[8 + (e & 1023)], row[e >> 10]); }
            bs = warp_sum(bs);
            __syncwarp();
            if (lane == 0) row[0] -= g * (1.0f - bs);          // a frame's occupancies sum to one: the blank's share
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_s2g(gb + (long long)t * p.sg_t, row, (uint32_t)V * 4u); bulk_commit(); }
        };
        if (NS >= 2) {
            // two stages: row k + 1 is gathered (its emissions are what the trellis warps wait for) before row k's
            // occupancies are awaited; row k + 2 is fetched into row k's stage once row k's store has read it
            for (int k = 0; k < min(2, nrows); ++k) issue(k);
            if (nrows > 0) pre(0);
            for (int k = 0; k < nrows; ++k) {
                if (k + 1 < nrows) pre(k + 1);
                post(k);
                if (k + 2 < nrows) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    issue(k + 2);
                }
            }
        } else {
            for (int k = 0; k < nrows; ++k) {
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
                issue(k);
                pre(k);
                post(k);
            }
        }
        if (lane == 0) bulk_wait_all<0>();
        return;
    }
    if (warp >= Wn) return;

    // ------------------------------------------------------------------------ trellis warps ---
    const int w = warp;
    const int nthr = 32 * Wn;
    const LaneCfg cfg = lane_cfg(32 * w + lane, dir, L, NL, p.NLmax, p.tgt + (size_t)n * p.Sp);
    Lane s;
    {
        const int* mybound = p.bound + ((size_t)n * 2 + dir) * p.BW;
        float4 ob = make_float4(0.f, 0.f, 0.f, 0.f), ol = ob;
        int4 oe = make_int4(kVoidE, kVoidE, kVoidE, kVoidE);
        if (cfg.live) {
            ob = *(const float4*)(mybound + 4 * cfg.g);
            ol = *(const float4*)(mybound + 4 * NL + 4 * cfg.g);
            oe = *(const int4*)(mybound + 8 * NL + 4 * cfg.g);
        }
        s.e[0] = oe.x; s.e[1] = oe.y; s.e[2] = oe.z; s.e[3] = oe.w;
        s.b[0] = ob.x; s.b[1] = ob.y; s.b[2] = ob.z; s.b[3] = ob.w;
        s.l[0] = ol.x; s.l[1] = ol.y; s.l[2] = ol.z; s.l[3] = ol.w;
    }
    const int4 zi = p.zinfo[n];
    const int eZ = zi.x;
    const float rZ = __int_as_float(zi.y);
    if (lane == 31) mail[((steps1 + 1) & 1) * W + w] = make_int2(__float_as_int(s.l[dir ? 0 : kJP - 1]), s.e[dir ? 0 : kJP - 1]);
    const int k_inj = dir ? 4 * NL - 1 - (L + 4) : 3;
    const int q_lo = 128 * w - k_inj - 1, q_hi = 128 * w + 127 - k_inj - 1;
    const int first = max(q_lo, 0), last = Tn - L + q_hi;
    float* emp = s_em + 4 + 4 * cfg.g;                      // my four label emissions / occupancies in slot 0
    const int* stp = s_st + 4 * cfg.g;                      // the other side's four label states of my group in slot 0
    const int eoff = 4 * NL;
    const uint32_t st_bytes = (uint32_t)(8 * NL) * 4u;
    const int* tr_n = p.tr + (size_t)n * p.T * p.SPL;
    auto st_issue = [&](int i2) {          // the row the OTHER side stored for the frame of my phase-2 step i2
        const int i = steps1 + i2, t = dir ? Tn - 1 - i : i;
        mbar_expect_tx(&st_full[i2 & (kNSR - 1)], st_bytes);
        bulk_g2s(s_st + (i2 & (kNSR - 1)) * p.SPL, tr_n + (size_t)t * p.SPL, st_bytes, &st_full[i2 & (kNSR - 1)]);
    };
    if (threadIdx.x == 0)
        for (int i2 = 0; i2 < min(kNSR - 1, nsteps2); ++i2) st_issue(i2);
    side_barrier(nthr);

    auto sweep = [&](auto dirc) {
        constexpr int DIR = decltype(dirc)::value;
        for (int i2 = 0; i2 < nsteps2; ++i2) {
            const int i = steps1 + i2;
            if (threadIdx.x == 0 && i2 + kNSR - 1 < nsteps2) st_issue(i2 + kNSR - 1);      // its slot was read in step i2 - 1
            const int slot = i2 & (kNE - 1), ss = i2 & (kNSR - 1);
            mbar_wait_sleep(&em_full[slot], (uint32_t)(i2 / kNE) & 1u, 32);
            const float pb = s_em[slot * EMF];
            const float4 pl = *(const float4*)(emp + slot * EMF);
            float4 occ = make_float4(0.f, 0.f, 0.f, 0.f);
            mbar_wait(&st_full[ss], (uint32_t)(i2 / kNSR) & 1u);
            if (i >= first && i <= last) {
                float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
                int4 eb = make_int4(kVoidE, kVoidE, kVoidE, kVoidE);
                if (cfg.live) { o4 = *(const float4*)(stp + ss * p.SPL); eb = *(const int4*)(stp + ss * p.SPL + eoff); }
                float cm; int ce;
                fetch_below<DIR>(s, w, lane, mail + ((i + 1) & 1) * W, cm, ce);
                float u[kJP], v[kJP];
                lane_sums<DIR>(s, cfg.allowed, cm, ce, u, v);
                // occupancy of a label state = (my pre-emission sum) x (the other side's stored value) / Z
                auto scale = [&](int xs) {       // 2^xs / mantissa of Z, flushing below 2^-126
                    return (xs < -126) ? 0.0f : __int_as_float(__float_as_int(rZ) + (min(xs, 90) << 23));
                };
                occ = make_float4((v[0] * o4.x) * scale(s.e[0] + eb.x - eZ), (v[1] * o4.y) * scale(s.e[1] + eb.y - eZ),
                                  (v[2] * o4.z) * scale(s.e[2] + eb.z - eZ), (v[3] * o4.w) * scale(s.e[3] + eb.w - eZ));
                lane_emit(s, u, v, pb, pl);
            }
            if (lane == 31) mail[(i & 1) * W + w] = make_int2(__float_as_int(s.l[DIR ? 0 : kJP - 1]), s.e[DIR ? 0 : kJP - 1]);
            if (cfg.live) *(float4*)(emp + slot * EMF) = occ;
            mbar_arrive(&occ_full[slot]);
            side_barrier(nthr);
        }
    };
    if (dir) sweep(std::integral_constant<int, 1>{}); else sweep(std::integral_constant<int, 0>{});
}

}  // namespace hab

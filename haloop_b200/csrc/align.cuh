// align.cuh — alignment paths: the reference's greedy argmax decode (ha/recognizer.py:48-59) and a
// max-semiring (Viterbi) forced alignment over the CTC trellis of ha/ctc.py:144-167.  Index and
// comparison work only: results are bit-exact against the CPU restatement the tests check with.
#pragma once
#include "common.cuh"
#include "ctc.cuh"

namespace hab {

// --------------------------------------------------------------------------- greedy argmax ---
struct GreedyParams {
    const float* x; long long sx_n, sx_t;
    int N, T, V;
    const void* in_len; int len64;
    long long* alignment; float* score; long long* hyp; long long* hyp_len;
};

// grid (ceil(T/8), N), block 256: one warp per frame; first index wins ties, as torch.max
__global__ void __launch_bounds__(256) greedy_argmax_kernel(GreedyParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y, t = blockIdx.x * 8 + warp;
    if (t >= p.T) return;
    const float* row = p.x + (long long)n * p.sx_n + (long long)t * p.sx_t;
    float best = -CUDART_INF_F;
    int arg = 0x7fffffff;
    for (int c = lane; c < p.V; c += 32) {
        const float v = row[c];
        if (v > best || arg == 0x7fffffff) { best = v; arg = c; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        p.alignment[(size_t)n * p.T + t] = arg;
        p.score[(size_t)n * p.T + t] = best;
    }
}

// grid N, block 256: unique_consecutive + drop blanks as a flag / prefix-sum / scatter
__global__ void __launch_bounds__(256) greedy_collapse_kernel(GreedyParams p) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int Tn = p.T;
    if (p.in_len) {
        long long v = load_idx(p.in_len, n, p.len64);
        Tn = v < 0 ? 0 : (v > p.T ? p.T : (int)v);
    }
    const long long* ali = p.alignment + (size_t)n * p.T;
    long long* hyp = p.hyp + (size_t)n * p.T;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int t0 = 0; t0 < p.T; t0 += 256) {
        const int t = t0 + tid;
        long long a = 0;
        int keep = 0;
        if (t < Tn) {
            a = ali[t];
            const long long prev = (t > 0) ? ali[t - 1] : -1;
            keep = (a != prev) && (a != 0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int within = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (keep) hyp[before + within] = a;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int m = s_base;
    for (int t = m + tid; t < p.T; t += 256) hyp[t] = -1;
    if (tid == 0) p.hyp_len[n] = m;
}

// ----------------------------------------------------------------------------- CTC Viterbi ---
struct ViterbiWs { size_t meta, order, tgt, dupnext, bp, total; int Sp, S_; };

__host__ inline ViterbiWs viterbi_ws_layout(int T, int N, int S) {
    ViterbiWs w;
    w.Sp = round_up(S > 0 ? S : 1, 4);
    w.S_ = 2 * S + 1;
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.dupnext = take(sizeof(int) * (size_t)N * w.Sp);
    w.bp = take((size_t)N * T * w.S_);
    w.total = o;
    return w;
}

struct ViterbiParams {
    const float* lp; long long sx_t, sx_n;
    int T, N, V, S_;
    const int4* meta; const int* tgt; int Sp;
    unsigned char* bp;
    long long* alignment; float* score;
};

// grid N, block 256, dynamic smem (2*S_ floats + S_ ints).  Candidates in the order (self, prev, skip);
// a later one wins only if strictly greater; float32 adds in time order, so the result is
// bit-reproducible on any IEEE machine.
__global__ void __launch_bounds__(256) ctc_viterbi_kernel(ViterbiParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int4 mt = p.meta[n];
    const int Tn = mt.z ? 0 : mt.x, L = mt.y;
    long long* ali = p.alignment + (size_t)n * p.T;
    for (int t = tid; t < p.T; t += blockDim.x) ali[t] = -1;
    if (Tn <= 0) { if (tid == 0) p.score[n] = mt.z ? CUDART_NAN_F : -CUDART_INF_F; return; }
    const int S_ = 2 * L + 1;
    float* v = (float*)smem_raw;                 // [2][S_]
    int* cls = (int*)(v + 2 * p.S_);
    const int* y = p.tgt + (size_t)n * p.Sp;
    for (int s = tid; s < S_; s += blockDim.x) cls[s] = (s & 1) ? (y[s >> 1] & kLabelMask) : 0;
    __syncthreads();
    const float* lpb = p.lp + (long long)n * p.sx_n;
    unsigned char* bp = p.bp + (size_t)n * p.T * p.S_;
    for (int s = tid; s < S_; s += blockDim.x) {
        float a = -CUDART_INF_F;
        if (s == 0) a = lpb[0];
        else if (s == 1) a = lpb[cls[1]];
        v[s] = a;
    }
    __syncthreads();
    int cur = 0;
    for (int t = 1; t < Tn; ++t) {
        const float* e = lpb + (long long)t * p.sx_t;
        const float* pv = v + cur * p.S_;
        float* nv = v + (cur ^ 1) * p.S_;
        for (int s = tid; s < S_; s += blockDim.x) {
            const int c = cls[s];
            float best = pv[s];
            unsigned char arg = 0;
            if (s >= 1 && pv[s - 1] > best) { best = pv[s - 1]; arg = 1; }
            if (s >= 2 && c != 0 && c != cls[s - 2] && pv[s - 2] > best) { best = pv[s - 2]; arg = 2; }
            nv[s] = __fadd_rn(best, e[c]);
            bp[(size_t)t * p.S_ + s] = arg;
        }
        cur ^= 1;
        __syncthreads();
    }
    if (tid == 0) {
        const float* fv = v + cur * p.S_;
        int s = S_ - 1;
        if (S_ > 1 && fv[S_ - 2] > fv[S_ - 1]) s = S_ - 2;
        p.score[n] = fv[s];
        __threadfence_block();
        if (!(fv[s] > -CUDART_INF_F)) return;        // no alignment exists: the score is -inf and the path stays all -1
        for (int t = Tn - 1; t >= 0; --t) {
            ali[t] = cls[s];
            if (t > 0) s -= bp[(size_t)t * p.S_ + s];
        }
    }
}

}  // namespace hab

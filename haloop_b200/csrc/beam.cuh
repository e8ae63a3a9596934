// beam.cuh — CTC prefix beam search on the GPU, hypothesis for hypothesis the reference's
// ctc_beam_search_decode_logits (ha/beam.py:71-137; "FIXME: speed it up", ha/recognizer.py:58).
//
// One CTA per utterance walks the frames; per frame the candidates are the current beams unchanged plus every beam
// extended by every class (class 0 included, as in the reference), scored in the log domain, and the `beam` best
// survive in descending order.  The reference's semantics are kept exactly where they are observable:
//   * beams are updated one after the other IN PLACE (ha/beam.py:99-111), so a beam whose prefix sits earlier in the
//     list sees that prefix's blank score of THIS frame, one whose prefix sits later sees last frame's;
//   * `top_seqs.index(seq[:-1])` takes the first beam equal to the prefix, and equal sequences reached along
//     different routes are NOT merged;
//   * extension candidates enter with a blank score of 0.0 (`torch.zeros` in the log domain, ha/beam.py:124), i.e.
//     log 1, not log 0: every extension therefore scores logaddexp(0, label score).  It is what the reference
//     computes, so it is what this kernel computes with ext_blank = 0.0f; ext_blank = -inf gives the algorithm the
//     reference's probability-domain twin (ha/beam.py:4-68, blank mass 0 for an extension) and [Graves14] describe.
// Sequences are compared through a 64-bit hash + length (kept per beam together with the hash of the sequence minus
// its last symbol) and reconstructed at the end from per-frame back-pointers.  Ties between candidate scores go to
// the lower candidate index (existing beams first, then extensions in (beam, class) order).
#pragma once
#include "common.cuh"

namespace hab {

constexpr int kMaxBeam = 16;

struct BeamParams {
    const float* lp; long long sx_n, sx_t;      // (N,T,V) log-probs, unit class stride
    int N, T, V, beam;
    float ext_blank;                             // blank score of an extension candidate: 0.0f = the reference, -inf = Graves
    const void* in_len; int len64;
    int* bp;                                     // [N][T][beam]: parent << 16 | (symbol + 1), symbol = -1: unchanged
    float* cand;                                 // [N][beam * (V + 1)] candidate scores
    long long* hyp; long long* hyp_len; float* score;    // (N,beam,T) padded with -1, (N,beam), (N,beam)
};

__host__ __device__ inline size_t beam_ws_bytes(int N, int T, int V, int beam) {
    return round_up_sz((size_t)N * T * beam * 4, 256) + round_up_sz((size_t)N * beam * (V + 1) * 4, 256);
}

__device__ __forceinline__ float logaddexp_f(float a, float b) {      // torch.logaddexp (fp32)
    if (isinf(a) && a == b) return a;
    const float m = fmaxf(a, b);
    return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ unsigned long long hash_push(unsigned long long h, int k) {
    h ^= (unsigned long long)(k + 1) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h * 0xff51afd7ed558ccdull;
}

// grid N, block 256
__global__ void __launch_bounds__(256) ctc_beam_kernel(BeamParams p) {
    __shared__ float s_seq[kMaxBeam], s_blank[kMaxBeam], s_label[kMaxBeam];
    __shared__ int s_len[kMaxBeam], s_last[kMaxBeam];
    __shared__ unsigned long long s_hash[kMaxBeam], s_phash[kMaxBeam];
    __shared__ float n_seq[kMaxBeam], n_blank[kMaxBeam], n_label[kMaxBeam];
    __shared__ int n_len[kMaxBeam], n_last[kMaxBeam], n_bp[kMaxBeam];
    __shared__ unsigned long long n_hash[kMaxBeam], n_phash[kMaxBeam];
    __shared__ float r_val[8]; __shared__ int r_idx[8];
    __shared__ int s_nb;
    const int n = blockIdx.x, tid = threadIdx.x, V = p.V, B = p.beam;
    long long Tn = p.in_len ? load_idx(p.in_len, n, p.len64) : p.T;
    if (Tn < 0) Tn = 0;
    if (Tn > p.T) Tn = p.T;
    const float* lpn = p.lp + (long long)n * p.sx_n;
    float* cand = p.cand + (size_t)n * B * (V + 1);
    int* bp = p.bp + (size_t)n * p.T * B;
    if (tid == 0) {
        s_nb = 1; s_seq[0] = 0.0f; s_blank[0] = 0.0f; s_label[0] = -CUDART_INF_F;
        s_len[0] = 0; s_last[0] = 0; s_hash[0] = 0x1234567ull; s_phash[0] = 0;
    }
    __syncthreads();
    for (int t = 0; t < (int)Tn; ++t) {
        const float* e = lpn + (long long)t * p.sx_t;
        const int nb = s_nb;
        if (tid == 0) {
            // ha/beam.py:99-111, beam after beam, in place
            for (int s = 0; s < nb; ++s) {
                if (s_len[s] > 0) {
                    const float el = e[s_last[s]];
                    s_label[s] += el;
                    for (int q = 0; q < nb; ++q)
                        if (s_len[q] == s_len[s] - 1 && s_hash[q] == s_phash[s]) {
                            s_label[s] = logaddexp_f(s_label[s], el + 0.0f + s_blank[q]);
                            break;
                        }
                }
                s_blank[s] = s_seq[s] + e[0];
            }
        }
        __syncthreads();
        // candidates: [0, nb) the beams unchanged, then (s, k) -> nb + s V + k
        const int ncand = nb + nb * V;
        for (int i = tid; i < ncand; i += 256) {
            float v;
            if (i < nb) v = logaddexp_f(s_blank[i], s_label[i]);
            else {
                const int s = (i - nb) / V, k = (i - nb) - s * V;
                const int last = s_len[s] > 0 ? s_last[s] : 0;
                const float ext = e[k] + 0.0f + ((k == last) ? s_blank[s] : s_seq[s]);
                v = logaddexp_f(p.ext_blank, ext);                      // blank component of an extension: torch.zeros (ha/beam.py:124)
            }
            cand[i] = v;
        }
        __syncthreads();
        const int nsel = min(B, ncand);
        for (int r = 0; r < nsel; ++r) {
            float bv = -CUDART_INF_F; int bi = 0x7fffffff;
            for (int i = tid; i < ncand; i += 256) {
                const float v = cand[i];
                if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if ((tid & 31) == 0) { r_val[tid >> 5] = bv; r_idx[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < 8; ++w)
                    if (r_val[w] > bv || (r_val[w] == bv && r_idx[w] < bi)) { bv = r_val[w]; bi = r_idx[w]; }
                // NaN scores never win a comparison: fall back to the first untaken candidate (bi stays in range)
                if (bi == 0x7fffffff) bi = 0;
                n_seq[r] = bv;
                if (bi < nb) {
                    n_blank[r] = s_blank[bi]; n_label[r] = s_label[bi]; n_len[r] = s_len[bi]; n_last[r] = s_last[bi];
                    n_hash[r] = s_hash[bi]; n_phash[r] = s_phash[bi]; n_bp[r] = (bi << 16) | 0;
                } else {
                    const int s = (bi - nb) / V, k = (bi - nb) - s * V;
                    const int last = s_len[s] > 0 ? s_last[s] : 0;
                    n_blank[r] = p.ext_blank;
                    n_label[r] = e[k] + 0.0f + ((k == last) ? s_blank[s] : s_seq[s]);
                    n_len[r] = s_len[s] + 1; n_last[r] = k;
                    n_phash[r] = s_hash[s]; n_hash[r] = hash_push(s_hash[s], k);
                    n_bp[r] = (s << 16) | (k + 1);
                }
                cand[bi] = __int_as_float(0x7fc00000);            // taken: a NaN never wins a comparison again
                r_idx[0] = bi;
            }
            __syncthreads();
        }
        if (tid < nsel) {
            s_seq[tid] = n_seq[tid]; s_blank[tid] = n_blank[tid]; s_label[tid] = n_label[tid];
            s_len[tid] = n_len[tid]; s_last[tid] = n_last[tid]; s_hash[tid] = n_hash[tid]; s_phash[tid] = n_phash[tid];
            bp[(size_t)t * B + tid] = n_bp[tid];
        }
        if (tid == 0) s_nb = nsel;
        __syncthreads();
    }
    // hypotheses in beam order, reconstructed from the back-pointers
    const int nb = s_nb;
    for (int b = tid; b < B; b += 256) {
        long long* h = p.hyp + ((size_t)n * B + b) * p.T;
        int len = 0;
        if (b < nb) {
            len = s_len[b];
            int cur = b, pos = len;
            for (int t = (int)Tn - 1; t >= 0 && pos > 0; --t) {
                const int w = bp[(size_t)t * B + cur];
                if (w & 0xffff) h[--pos] = (w & 0xffff) - 1;
                cur = w >> 16;
            }
        }
        for (int i = len; i < p.T; ++i) h[i] = -1;
        p.hyp_len[(size_t)n * B + b] = (b < nb) ? len : 0;
        p.score[(size_t)n * B + b] = (b < nb) ? s_seq[b] : -CUDART_INF_F;
    }
}

}  // namespace hab

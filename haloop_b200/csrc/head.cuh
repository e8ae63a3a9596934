// head.cuh — fused classifier head + CTC for sm_100a: replaces
//     features.log_softmax(classifier(dropout(features)))  ->  ctc loss      ha/recognizer.py:43-46, 61-82
// without ever writing the (N,T,V) logits, log-probs or their gradient to HBM.
//
// One tensor-core GEMM kernel (umma_gemm_kernel, umma_gemm.cuh) serves the four contractions of the op, told apart by its epilogue:
//   kEpiFwd    S = h W^T + b   per 128 x 128 tile: row max / sum-exp partials and the blank + label logits of the
//                              row's utterance gathered into the CTC emission rows (logits never leave the SM)
//   kEpiBwdD   S again (recomputed), d = g (softmax(S) - occupancy): written for ONE chunk of rows (an L2-sized
//                              ring, not an (N,T,V) gradient)
//   kEpiStore  dh = d W        W read as it lies (MN-major B operand); plain store
//   kEpiAccum  dW += d^T h     both operands read as they lie (MN-major: the contraction runs over their rows);
//                              split-K partial accumulators, one writer per element (deterministic)
//
// The GEMM engine (TMA operand ring, tf32 x 3 split in shared memory, chunked TMEM accumulation) is umma_gemm.cuh.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace hab {

enum { kEpiFwd = 0, kEpiBwdD = 1, kEpiStore = 2, kEpiAccum = 3 };
constexpr int kHasDup = 0x40000000;           // cls2pos flag: the class occurs again later in the target

struct HeadGemmParams {
    int M, N, K;                 // rows of A this launch covers, rows of B (= output columns), contraction length
    int a_row0;                  // added to A's row coordinate (start of the row chunk)
    int splits;                  // split-K request (dW)
    int a_mn, b_mn, a_k0, b_k0;  // MN-major operands and their K-coordinate offsets (GemmCore)
    int Mpad;                    // kEpiBwdD: rows of `out` this launch may write (M rounded up to the tile: zeros beyond M)
    int nprod;                   // 3: error-compensated tf32 x 3; 1: plain tf32
    // epilogue data
    const float* bias;           // (N) or null
    int rows_total, T;           // all rows of h (= Nutt * T); frames per utterance
    const int4* meta; const int* cls2pos; const int* dupnext; int Sp, V;
    float* em; int E;            // emission rows (fwd: raw logits gathered in; bwd: occupancy rows)
    float2* stats;               // fwd: [tiles_n][rows_total] (max, sum exp) partials
    const float* lse2; const float* loss; const float* gout;
    float* out; long long ldo;   // kEpiBwdD: d (M x N chunk), kEpiStore: dh, kEpiAccum: partials [splits][M][N]
    int accumulate;              // kEpiAccum: add to what is there (every chunk but the first)
};

// Epilogues of the four head contractions (the code after the accumulator has been folded into SUM; `trow` addresses
// this thread's TMEM lane of SUM).  mb / nb: tile coordinates, sp: split-K index, q: lane quarter of the warp.
template <int EPI>
struct HeadEpi {
    HeadGemmParams p;
    __device__ __forceinline__ void operator()(uint32_t trow, int mb, int nb, int sp, int /*batch*/, int q, int lane, float* stg) const {
            const int rl = mb * kHM + 32 * q + lane;             // row inside this launch's A range
        const int n0 = nb * kHN;

        if (EPI == kEpiFwd || EPI == kEpiBwdD) {
            const int row = p.a_row0 + rl;                   // global row of h = utterance * T + frame
            const bool inside = rl < p.M && row < p.rows_total;
            const int n = inside ? row / p.T : 0, t = row - n * p.T;
            const int4 mt = p.meta[n];
            bool live = inside && !mt.z && t < mt.x;
            float g = 0.0f, l2 = 0.0f;
            if (EPI == kEpiBwdD) {
                const float lossn = p.loss[n];
                live = live && (lossn < CUDART_INF_F);
                if (live) { g = p.gout[n]; l2 = p.lse2[row]; }
            }
            const int* c2p = p.cls2pos + (size_t)n * p.V;
            const int* nxt = p.dupnext + (size_t)n * p.Sp;
            float* erow = p.em + (size_t)row * p.E;
            float m = -CUDART_INF_F, ssum = 0.0f;
            for (int c0 = 0; c0 < kHN && n0 + c0 < p.N; c0 += 16) {
                float a[16];
                tmem_ld16_nowait(trow + c0, a);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = n0 + c0 + j;
                    const float bj = (p.bias && col < p.N) ? __ldg(p.bias + col) : 0.0f;
                    v[j] = a[j] + bj;
                }
                if (EPI == kEpiFwd) {
                    float cm = -CUDART_INF_F;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n0 + c0 + j < p.N) cm = fmaxf(cm, v[j]);
                    const float mn = fmaxf(m, cm);
                    float acc = 0.0f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (n0 + c0 + j < p.N) acc += ex2f((v[j] - mn) * kLog2e);
                    ssum = ssum * ex2f((m - mn) * kLog2e) + acc;
                    m = mn;
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = n0 + c0 + j;
                            if (col < p.N) {
                                if (col == 0) erow[1] = v[j];
                                const int kf = __ldg(c2p + col);
                                if (kf >= 0) {
                                    erow[4 + (kf & ~kHasDup)] = v[j];
                                    if (kf & kHasDup)
                                        for (int k = __ldg(nxt + (kf & ~kHasDup)); k >= 0; k = __ldg(nxt + k)) erow[4 + k] = v[j];
                                }
                            }
                        }
                    }
                } else {
                    // occupancy of class col = sum over the positions that hold it.  All sixteen first-position loads of
                    // the chunk are issued before any is used (they are scattered 4-byte reads of this thread's own
                    // occupancy row: one at a time they cost an L2 round trip each); later occurrences are rare
                    int kf[16];
                    float o[16], dv[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) kf[j] = (live && n0 + c0 + j < p.N) ? __ldg(c2p + n0 + c0 + j) : -1;
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = (kf[j] >= 0) ? erow[4 + (kf[j] & ~kHasDup)] : 0.0f;
                    if (live && n0 + c0 == 0) o[0] += 1.0f - erow[1];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (kf[j] >= 0 && (kf[j] & kHasDup))
                            for (int k = __ldg(nxt + (kf[j] & ~kHasDup)); k >= 0; k = __ldg(nxt + k)) o[j] += erow[4 + k];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        dv[j] = (live && n0 + c0 + j < p.N) ? g * (ex2f(fmaf(v[j], kLog2e, -l2)) - o[j]) : 0.0f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = dv[j];
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + (lane >> 2), cq = (lane & 3) * 4;
                        const int mm = mb * kHM + 32 * q + r, c = n0 + c0 + cq;
                        if (mm < p.Mpad && c < p.N) {
                            const float* a4 = stg + r * 17 + cq;
                            *(float4*)(p.out + (size_t)mm * p.ldo + c) = make_float4(a4[0], a4[1], a4[2], a4[3]);
                        }
                    }
                    __syncwarp();
                }
            }
            if (EPI == kEpiFwd && inside) p.stats[(size_t)nb * p.rows_total + row] = make_float2(m, ssum);
        } else {
            // thread = accumulator row (fixed by the TMEM lane mapping), but a row-per-thread store touches 32 rows of 16
            // bytes per instruction: the 32 x 16 block of a warp goes through its staging block in shared memory and
            // leaves (or is read-modify-written) as 64-byte row segments, 8 rows per instruction
            float* obase = (EPI == kEpiAccum) ? p.out + (size_t)sp * p.M * p.ldo : p.out + (size_t)p.a_row0 * p.ldo;
            const int m_w = mb * kHM + 32 * q;
            for (int c0 = 0; c0 < kHN && n0 + c0 < p.N; c0 += 16) {
                float a[16];
                tmem_ld16_nowait(trow + c0, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = a[j];
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + (lane >> 2), cq = (lane & 3) * 4;
                    const int mm = m_w + r, c = n0 + c0 + cq;
                    if (mm < p.M && c < p.N && (EPI == kEpiAccum || p.a_row0 + mm < p.rows_total)) {
                        const float* a4 = stg + r * 17 + cq;
                        float4 o = make_float4(a4[0], a4[1], a4[2], a4[3]);
                        float4* dst = (float4*)(obase + (size_t)mm * p.ldo + c);
                        if (EPI == kEpiAccum && p.accumulate) {
                            const float4 old = *dst;
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *dst = o;
                    }
                }
                __syncwarp();
            }
        }
    }
};

// ------------------------------------------------------------------------------- small kernels ---
// grid N, block 256: cls2pos[n][c] = first target position of utterance n that holds class c (| kHasDup when a later
// position holds it too: only then is the dupnext chain walked), else -1
__global__ void __launch_bounds__(256) head_cls2pos_kernel(const int4* meta, const int* tgt, const int* dupnext, int Sp, int V, int* cls2pos) {
    const int n = blockIdx.x;
    int* row = cls2pos + (size_t)n * V;
    for (int c = threadIdx.x; c < V; c += 256) row[c] = -1;
    __syncthreads();
    const int4 mt = meta[n];
    const int L = mt.z ? 0 : mt.y;
    for (int k = threadIdx.x; k < L; k += 256) {
        const int w = tgt[(size_t)n * Sp + k];
        if (!(w & kNotFirst)) row[w & kLabelMask] = k | (dupnext[(size_t)n * Sp + k] >= 0 ? kHasDup : 0);
    }
}

// grid ceil(rows / 8), block 256 (warp per row): combine the per-tile (max, sum exp) partials into the row's
// log2-sum-exp2 and turn the raw logits gathered into the emission row into CTC emissions (as ctc_rows_kernel does)
struct HeadFinalizeParams {
    int rows_total, T, tiles_n;
    const int4* meta; const float2* stats;
    float* lse2; float* em; int E;
};
__global__ void __launch_bounds__(256) head_finalize_kernel(HeadFinalizeParams p) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.rows_total) return;
    const int n = row / p.T, t = row - n * p.T;
    const int4 mt = p.meta[n];
    if (mt.z || t >= mt.x) return;
    float m = -CUDART_INF_F;
    for (int j = lane; j < p.tiles_n; j += 32) m = fmaxf(m, p.stats[(size_t)j * p.rows_total + row].x);
    m = warp_max(m);
    float s = 0.0f;
    for (int j = lane; j < p.tiles_n; j += 32) {
        const float2 st = p.stats[(size_t)j * p.rows_total + row];
        s += st.y * ex2f((st.x - m) * kLog2e);
    }
    s = warp_sum(s);
    const float l2 = fmaf(m, kLog2e, log2f(s));
    const float ct = round_int(fmaf(m, kLog2e, -l2));
    float* erow = p.em + (size_t)row * p.E;
    const int L = mt.y;
    for (int k = lane; k < L; k += 32) {
        float K, f;
        emission_split(erow[4 + k], l2, ct, K, f);
        erow[4 + k] = emission_linear(K, f);
    }
    if (lane == 0) {
        p.lse2[row] = l2;
        float K, f;
        emission_split(erow[1], l2, ct, K, f);
        *(float4*)erow = make_float4(ct, emission_linear(K, f), 0.0f, 0.0f);
    }
}

// bias gradient: column sums of the chunk's d (rows x V, row-major).  grid (ceil(V / 32), kColSplits), block (32, 8):
// a block sums its slice of the rows for 32 classes (coalesced 128-byte reads) into part[split][V]; the finish kernel
// adds the splits in a fixed order (deterministic) onto db.
constexpr int kColSplits = 32;
__global__ void __launch_bounds__(256) head_colsum_kernel(const float* d, int ldd, int rows, int V, float* part) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int per = (rows + kColSplits - 1) / kColSplits;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float s = 0.0f;
    if (c < V)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += d[(size_t)r * ldd + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < V) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        part[(size_t)blockIdx.y * V + c] = t;
    }
}
__global__ void __launch_bounds__(256) head_colsum_finish_kernel(const float* part, int V, float* db, int accumulate) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= V) return;
    float s = 0.0f;
    for (int k = 0; k < kColSplits; ++k) s += part[(size_t)k * V + c];
    db[c] = accumulate ? db[c] + s : s;
}

// dW[i] = sum over the split-K partials
__global__ void __launch_bounds__(256) head_reduce_kernel(const float* part, int splits, size_t n, float* out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float s = 0.0f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    out[i] = s;
}

}  // namespace hab

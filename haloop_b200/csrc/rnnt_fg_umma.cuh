// rnnt_fg_umma.cuh — the three contractions of the joint-free RNN-T loss (rnnt_fg.cuh) on the 5th-generation
// tensor cores, through the GEMM engine of umma_gemm.cuh (TMA operand ring, tcgen05.mma kind::tf32 with the
// error-compensated three-product split formed in shared memory, chunked round-to-nearest accumulation in TMEM):
//
//   E  = F G^T            (T x V)(V x U1)     K = V      -> arc probabilities for the lattice kernel
//   DF = (W G) (.) F      (T x U1)(U1 x V)    K = U1     -> d loss / d f
//   DG = (W^T F) (.) G    (U1 x T)(T x V)     K = T      -> d loss / d g
//
// Operands are exponentials of shifted logits (F, G) or occupancy ratios (W): small row kernels form them once as
// plain fp32 (kind::tf32 reads the high half of an fp32 word by itself; the engine's split warps derive the low half),
// zero-padded to whole tiles, and used as they lie: K-major where the contraction runs over their columns, MN-major
// where it runs over their rows (no transposed copies); every utterance is one batch entry of the engine.  A gradient needs 1e-5 absolute, so the accumulation chunks are short
// (16 k-blocks = 64 accumulating MMAs between round-to-nearest folds).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace hab {

constexpr int kUM = 128;          // accumulator rows per CTA (UMMA M)
constexpr int kUK = 16;           // k elements per stage (two tf32 UMMA k-steps of 8)
constexpr int kFgChunkKb = 16;    // k-blocks per accumulation chunk of the engine

struct FgUmmaWs {                 // operand buffers (byte offsets from the start of this block); floats per utterance
    size_t Fh, Fl, Gh, Gl, Fth, Ftl, Gth, Gtl, Wh, Wl, Wth, Wtl, total;
    int Tp, Tk, Uk, Um;           // T padded to 128 / 16, U1 padded to 16 / 128
};
__host__ inline FgUmmaWs fg_umma_ws_layout(int N, int T, int U1, int V) {
    FgUmmaWs w;
    w.Tp = round_up(T, kUM); w.Tk = round_up(T, kUK);
    w.Uk = round_up(U1, kUK);                                       // columns of E / contraction length of DF
    w.Um = max(round_up(U1, kUM), w.Uk);                            // rows of the G buffers (E reads Uk of them) and DG's M extent
    size_t o = 0;
    auto take = [&](size_t floats) { size_t at = o; o = round_up_sz(o + floats * 4 * (size_t)N, 256); return at; };
    w.Fh = take((size_t)w.Tp * V); w.Fl = take(0);                         // the low halves are formed in shared memory
    w.Gh = take((size_t)w.Um * V); w.Gl = take(0);                         // Um rows: G is also the M operand's epilogue source
    w.Fth = take(0); w.Ftl = take(0);                                       // no transposed copies: MN-major operands
    w.Gth = take(0); w.Gtl = take(0);
    w.Wh = take((size_t)w.Tp * w.Uk); w.Wl = take(0);
    w.Wth = take(0); w.Wtl = take(0);
    w.total = o;
    return w;
}

// ---------------------------------------------------------------------------------- row kernels ---
struct FgUmmaParams {
    const float* f; const float* g; float* gf; float* gg;
    int N, T, U1, V, Up, D;
    const int4* meta; const int* tgt;
    float* mf; float* mg; float* lf0; float* lg0; float* lgy; float* E;
    float2* bl; float2* lb; const float2* occ;
    const float* gout; const float* loss;
    float *Fh, *Fl, *Gh, *Gl, *Fth, *Ftl, *Gth, *Gtl, *Wh, *Wl, *Wth, *Wtl;
    int Tp, Tk, Uk, Um;
};

// grid (ceil((Tp + Um) / 8), N), block 256: one warp per (padded) row of f and of g.  Row statistics (as
// rnnt_fg_stats_kernel) and the operands F = exp(f - max f), G = exp(g - max g); padding rows are zero.
__global__ void __launch_bounds__(256) fg_rows_kernel(FgUmmaParams p) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= p.Tp + p.Um) return;
    const int4 mt = p.meta[n];
    const int Tn = mt.z ? 0 : mt.x, Un = mt.y;
    const bool isg = r >= p.Tp;
    const int row = isg ? r - p.Tp : r;
    const int V = p.V;
    float* oh = isg ? p.Gh + ((size_t)n * p.Um + row) * V : p.Fh + ((size_t)n * p.Tp + row) * V;
    const bool inside = isg ? row < p.U1 : row < p.T;              // a row of the input tensor (statistics are kept for all)
    const bool valid = isg ? (row <= Un && Tn > 0) : row < Tn;     // a row the loss uses
    if (!inside) {
        for (int c = lane; c < V; c += 32) oh[c] = 0.0f;
        return;
    }
    const float* x = isg ? p.g + ((size_t)n * p.U1 + row) * V : p.f + ((size_t)n * p.T + row) * V;
    float mx = -CUDART_INF_F;
    for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    const float m2 = mx * kLog2e;
    if (lane == 0) {
        const float l0 = fmaf(x[0], kLog2e, -m2);
        if (isg) {
            p.mg[(size_t)n * p.U1 + row] = m2;
            p.lg0[(size_t)n * p.U1 + row] = l0;
            const int y = (row < mt.y) ? (p.tgt[(size_t)n * p.Up + row] & kLabelMask) : 0;
            p.lgy[(size_t)n * p.U1 + row] = fmaf(x[y], kLog2e, -m2);
        } else {
            p.mf[(size_t)n * p.T + row] = m2;
            p.lf0[(size_t)n * p.T + row] = l0;
        }
    }
    for (int c = lane; c < V; c += 32) {
        oh[c] = valid ? ex2f(fmaf(x[c], kLog2e, -m2)) : 0.0f;
    }
}

// grid (ceil(Tp / 8), N), block 256 (one warp per padded frame): W[t][u] = (occ_blank + occ_label)[t][u] / E[t][u],
// as W (Tp x Uk); zero outside the utterance
__global__ void __launch_bounds__(256) fg_w_kernel(FgUmmaParams p) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    if (t >= p.Tp) return;
    const int4 mt = p.meta[n];
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x, Un = mt.y;
    const float2* occ = p.occ + (size_t)n * p.D * p.U1;
    const float* Em = p.E + (size_t)n * p.T * p.U1;
    for (int u = lane; u < p.Uk; u += 32) {
        float w = 0.0f;
        if (t < Tn && u <= Un) {
            const float2 o = occ[(size_t)(t + u) * p.U1 + u];
            const float e = Em[(size_t)t * p.U1 + u];
            if (e > 0.0f) w = (o.x + o.y) / e;
        }
        p.Wh[((size_t)n * p.Tp + t) * p.Uk + u] = w;
    }
}

// grid (ceil(T / 8), N), block 256 (one warp per frame, lanes over u): blank / label arc probabilities of every lattice
// node from E = F G^T, written diagonal-major for the lattice kernel
//     blank(t,u) = f[t][0] + g[u][0] - lse(t,u),  label(t,u) = f[t][y_u] + g[u][y_u] - lse(t,u),  lse = mf + mg + log2 E
__global__ void __launch_bounds__(256) fg_arc_kernel(FgUmmaParams q) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    const int4 mt = q.meta[n];
    const int Tn = mt.z ? 0 : mt.x, Un = mt.y;
    if (t >= Tn) return;
    const int T = q.T, U1 = q.U1, V = q.V;
    const int* y = q.tgt + (size_t)n * q.Up;
    const float* f = q.f + ((size_t)n * T + t) * V;
    const float* Er = q.E + ((size_t)n * T + t) * U1;
    float2* bl = q.bl + (size_t)n * q.D * U1;
    float2* lb = q.lb + (size_t)n * q.D * U1;
    const float lf0 = q.lf0[(size_t)n * T + t], mft = q.mf[(size_t)n * T + t];
    for (int u = lane; u <= Un; u += 32) {
        const float lE = log2f(Er[u]);
        const size_t sk = (size_t)(t + u) * U1 + u;
        bl[sk] = log2_to_parts(lf0 + q.lg0[(size_t)n * U1 + u] - lE);
        float ll = kVoid;
        if (u < Un) ll = fmaf(f[y[u] & kLabelMask], kLog2e, -mft) + q.lgy[(size_t)n * U1 + u] - lE;
        lb[sk] = log2_to_parts(ll);
    }
}

// ---------------------------------------------------------------------------- the three epilogues ---
enum { kUmmaE = 0, kUmmaDF = 1, kUmmaDG = 2 };

// Epilogue functor for umma_gemm_kernel: batch entry n = utterance, (mb, nb) = 128 x 128 tile of E, DF or DG; `trow`
// addresses this thread's TMEM lane of the folded accumulator (row m of the tile = lane 32 q + lane).
template <int MODE>
struct FgEpi {
    FgUmmaParams q;
    __device__ __forceinline__ void operator()(uint32_t trow, int mb, int nb, int /*sp*/, int n, int qq, int lane, float* stg) const {
        const int m0 = mb * kHM, n0 = nb * kHN;
        const int4 mt = q.meta[n];
        const float lossn = (MODE == kUmmaE) ? 0.0f : q.loss[n];
        const int Tn = (mt.z || (MODE != kUmmaE && !(lossn < CUDART_INF_F))) ? 0 : mt.x, Un = mt.y;
        const int T = q.T, U1 = q.U1, V = q.V;
        if (MODE == kUmmaE) {
            // E[t][u] only (staged so that a warp stores 64-byte row segments); the arc probabilities are formed from it
            // by fg_arc_kernel with one thread per lattice node (done here, a thread would walk its row's U1 nodes - each
            // with a scattered label-logit load - one after the other, and a CTA owns a single tile of this contraction)
            float* Eo = q.E + (size_t)n * T * U1;
            const int m_w = m0 + 32 * qq;
            for (int c0 = 0; c0 < kHN && n0 + c0 < q.Uk; c0 += 16) {
                float acc[16];
                tmem_ld16(trow + c0, acc);
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = acc[j];
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int r = it * 16 + (lane >> 1), u0 = n0 + c0 + (lane & 1) * 8;
                    const int t = m_w + r;
                    if (t < Tn)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (u0 + j <= Un) Eo[(size_t)t * U1 + u0 + j] = stg[r * 17 + (lane & 1) * 8 + j];
                }
                __syncwarp();
            }
        } else {
            const float go = q.gout[n];
            const int M = (MODE == kUmmaDF) ? T : U1;
            float* out = (MODE == kUmmaDF) ? q.gf + (size_t)n * T * V : q.gg + (size_t)n * U1 * V;
            // thread = accumulator row (fixed by the TMEM lane mapping), but a row-per-thread store would touch 32 rows
            // of 16 bytes per instruction: the 32 x 16 block of a warp goes through its staging block in shared memory
            // and leaves as 64-byte row segments, 8 rows per instruction, multiplied by the factor F[m][c] (or G[m][c])
            const int m_w = m0 + 32 * qq;
            const float* X = (MODE == kUmmaDF) ? q.Fh : q.Gh;
            const size_t xrows = (MODE == kUmmaDF) ? (size_t)n * q.Tp : (size_t)n * q.Um;
            // the factor values of the NEXT 16-column chunk are requested before this chunk's accumulators are read, so
            // the global-load latency hides behind the TMEM load, the staging pass and the stores (the contraction
            // itself is short here - K = U1 or T - and the tile is bound by this epilogue)
            auto load_x = [&](int c0, float4 (&h)[4]) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int mm = m_w + it * 8 + (lane >> 2), c = n0 + c0 + (lane & 3) * 4;
                    const bool lv = (MODE == kUmmaDF) ? (mm < Tn) : (mm <= Un && Tn > 0);
                    h[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lv && mm < M && c < V && c0 < kHN) h[it] = __ldg((const float4*)(X + (xrows + mm) * V + c));
                }
            };
            auto emit = [&](int c0, const float4 (&h)[4]) {
                float acc[16];
                tmem_ld16(trow + c0, acc);
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = acc[j];
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + (lane >> 2), cq = (lane & 3) * 4;
                    const int mm = m_w + r, c = n0 + c0 + cq;
                    if (mm < M && c < V) {
                        const float* a4 = stg + r * 17 + cq;
                        *(float4*)(out + (size_t)mm * V + c) =
                            make_float4(go * a4[0] * h[it].x, go * a4[1] * h[it].y, go * a4[2] * h[it].z, go * a4[3] * h[it].w);
                    }
                }
                __syncwarp();
            };
            float4 hA[4], hB[4];
            load_x(0, hA);
            for (int c0 = 0; c0 < kHN && n0 + c0 < V; c0 += 32) {
                load_x(c0 + 16, hB);
                emit(c0, hA);
                if (c0 + 16 < kHN && n0 + c0 + 16 < V) {
                    load_x(c0 + 32, hA);
                    emit(c0 + 16, hB);
                }
            }
        }
    }
};

}  // namespace hab

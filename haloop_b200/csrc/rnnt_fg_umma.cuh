// rnnt_fg_umma.cuh — the three contractions of the joint-free RNN-T loss (rnnt_fg.cuh) on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32 with fp32 accumulators in tensor memory, error-compensated by a
// three-product split so that the result keeps ~21 significant bits (a 1e-5 gradient needs more than tf32's 10):
//
//     a = a_hi + a_lo,  a_hi = a with the 13 low mantissa bits cleared (exactly what kind::tf32 reads),
//                       a_lo = (a - a_hi) likewise                      [a - a_hi is exact in fp32]
//     a b  ~=  a_hi b_hi + a_hi b_lo + a_lo b_hi                         [the dropped a_lo b_lo is 2^-22 relative]
//
//   E  = F G^T            (T x V)(V x U1)     K = V      -> arc probabilities for the lattice kernel
//   DF = (W G) (.) F      (T x U1)(U1 x V)    K = U1     -> d loss / d f
//   DG = (W^T F) (.) G    (U1 x T)(T x V)     K = T      -> d loss / d g
//
// Operands are exponentials of shifted logits (F, G) or occupancy ratios (W): small row kernels form them once,
// split them, and store both halves K-major (and transposed where the contraction runs over the other index) in
// the workspace, zero-padded to whole tiles.  One GEMM kernel serves all three: 128 threads copy 128 x 16 operand
// tiles into shared memory in the canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core
// matrix), one elected thread issues the six tcgen05.mma of a stage (2 k-steps x 3 split products) and commits
// them to an mbarrier that frees the stage; the accumulator (128 lanes x N columns of TMEM) is read back with
// tcgen05.ld for the epilogue of each contraction.
//
// The tensor core adds into its fp32 accumulator by TRUNCATION, so a long chain of accumulating MMAs drifts by
// about half an ulp per instruction (measured: 2.5e-5 .. 5.8e-5 on these gradients with one accumulator over the
// whole K).  The kernel therefore keeps three TMEM regions: MAIN takes the a_hi b_hi products of at most kChunk
// stages (8 accumulating instructions), SMALL the two cross products (2^-11 of the magnitude: their drift does not
// matter), and after every chunk the four warps add MAIN + SMALL into SUM with round-to-nearest fp32 adds
// (tcgen05.ld / tcgen05.st).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hab {

constexpr int kUM = 128;          // accumulator rows per CTA (UMMA M)
constexpr int kUK = 16;           // k elements per stage (two tf32 UMMA k-steps of 8)
constexpr int kChunk = 4;         // stages per accumulation chunk (see the note on truncating accumulation)

struct FgUmmaWs {                 // operand buffers (byte offsets from the start of this block); floats per utterance
    size_t Fh, Fl, Gh, Gl, Fth, Ftl, Gth, Gtl, Wh, Wl, Wth, Wtl, total;
    int Tp, Tk, Uk, Um;           // T padded to 128 / 16, U1 padded to 16 / 128
};
__host__ inline FgUmmaWs fg_umma_ws_layout(int N, int T, int U1, int V) {
    FgUmmaWs w;
    w.Tp = round_up(T, kUM); w.Tk = round_up(T, kUK);
    w.Uk = (U1 <= 160) ? round_up(U1, kUK) : round_up(U1, 128);     // E's accumulator tiles: Uk columns, or 128 when U1 > 160
                                                                    // (three regions of a tile must fit the 512 TMEM columns)
    w.Um = max(round_up(U1, kUM), w.Uk);                            // rows of the G buffers (E reads Uk of them) and DG's M extent
    size_t o = 0;
    auto take = [&](size_t floats) { size_t at = o; o = round_up_sz(o + floats * 4 * (size_t)N, 256); return at; };
    w.Fh = take((size_t)w.Tp * V); w.Fl = take((size_t)w.Tp * V);
    w.Gh = take((size_t)w.Um * V); w.Gl = take((size_t)w.Um * V);          // Um rows: G is also the M operand's epilogue source
    w.Fth = take((size_t)V * w.Tk); w.Ftl = take((size_t)V * w.Tk);
    w.Gth = take((size_t)V * w.Uk); w.Gtl = take((size_t)V * w.Uk);
    w.Wh = take((size_t)w.Tp * w.Uk); w.Wl = take((size_t)w.Tp * w.Uk);
    w.Wth = take((size_t)w.Um * w.Tk); w.Wtl = take((size_t)w.Um * w.Tk);
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
// rnnt_fg_stats_kernel) and the split operands F = exp(f - max f), G = exp(g - max g); padding rows are zero.
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
    float* ol = isg ? p.Gl + ((size_t)n * p.Um + row) * V : p.Fl + ((size_t)n * p.Tp + row) * V;
    const bool inside = isg ? row < p.U1 : row < p.T;              // a row of the input tensor (statistics are kept for all)
    const bool valid = isg ? (row <= Un && Tn > 0) : row < Tn;     // a row the loss uses
    if (!inside) {
        for (int c = lane; c < V; c += 32) { oh[c] = 0.0f; ol[c] = 0.0f; }
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
        float hi = 0.0f, lo = 0.0f;
        if (valid) tf32_split(ex2f(fmaf(x[c], kLog2e, -m2)), hi, lo);
        oh[c] = hi; ol[c] = lo;
    }
}

// grid (ceil(C / 32), ceil(Rd / 32), 2 N), block (32, 8): dst[c][r] = src[r][c] for r < Rs (else 0), r < Rd
// (blockIdx.z & 1 selects the hi / lo array of the pair)
struct FgTransposeParams { const float* sh; const float* sl; float* dh; float* dl; int Rs, C, Rd, lds_rows; };
__global__ void __launch_bounds__(256) fg_transpose_kernel(FgTransposeParams p) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z >> 1;
    const float* src = ((blockIdx.z & 1) ? p.sl : p.sh) + (size_t)n * p.lds_rows * p.C;
    float* dst = ((blockIdx.z & 1) ? p.dl : p.dh) + (size_t)n * p.C * p.Rd;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < p.Rs && c < p.C) ? src[(size_t)r * p.C + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < p.C && r < p.Rd) dst[(size_t)c * p.Rd + r] = tile[threadIdx.x][i];
    }
}

// grid (ceil(Tp / 8), N), block 256 (one warp per padded frame): W[t][u] = (occ_blank + occ_label)[t][u] / E[t][u],
// split, as W (Tp x Uk) and W^T (Um x Tk); zero outside the utterance
__global__ void __launch_bounds__(256) fg_w_kernel(FgUmmaParams p) {
    const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + warp;
    if (t >= p.Tp) return;
    const int4 mt = p.meta[n];
    const float lossn = p.loss[n];
    const int Tn = (mt.z || !(lossn < CUDART_INF_F)) ? 0 : mt.x, Un = mt.y;
    const float2* occ = p.occ + (size_t)n * p.D * p.U1;
    const float* Em = p.E + (size_t)n * p.T * p.U1;
    for (int u = lane; u < p.Um; u += 32) {
        float hi = 0.0f, lo = 0.0f;
        if (t < Tn && u <= Un) {
            const float2 o = occ[(size_t)(t + u) * p.U1 + u];
            const float e = Em[(size_t)t * p.U1 + u];
            if (e > 0.0f) tf32_split((o.x + o.y) / e, hi, lo);
        }
        if (u < p.Uk) { p.Wh[((size_t)n * p.Tp + t) * p.Uk + u] = hi; p.Wl[((size_t)n * p.Tp + t) * p.Uk + u] = lo; }
        if (t < p.Tk) { p.Wth[((size_t)n * p.Um + u) * p.Tk + t] = hi; p.Wtl[((size_t)n * p.Um + u) * p.Tk + t] = lo; }
    }
}

// ------------------------------------------------------------------------------------ the GEMM ---
enum { kUmmaE = 0, kUmmaDF = 1, kUmmaDG = 2 };

struct FgGemmParams {
    const float* Ah; const float* Al; int lda; size_t a_batch;      // A: (rows x K) K-major, hi / lo, per-utterance stride
    const float* Bh; const float* Bl; int ldb; size_t b_batch;      // B: (cols x K) K-major
    int K;                                                           // multiple of kUK
    int NT;                                                          // accumulator columns of a CTA (multiple of 16, <= 256)
    FgUmmaParams q;
};

constexpr int kUStages = 4;       // operand stage buffers: two stages in flight, one being multiplied, one draining

// shared memory: [kUStages][A hi | A lo | B hi | B lo] tiles in core-matrix layout, then the mbarriers and the TMEM address
__host__ __device__ inline size_t fg_gemm_smem(int NT) { return kUStages * (size_t)(2 * kUM + 2 * NT) * kUK * 4 + 64; }

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid (M / 128, ceil(Ncols / NT), N utterances), block 128.
template <int MODE>
__global__ void __launch_bounds__(128) fg_umma_gemm_kernel(FgGemmParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z, m0 = blockIdx.x * kUM, n0 = blockIdx.y * p.NT;
    const FgUmmaParams& q = p.q;
    const int4 mt = q.meta[n];
    const float lossn = (MODE == kUmmaE) ? 0.0f : q.loss[n];
    const int Tn = (mt.z || (MODE != kUmmaE && !(lossn < CUDART_INF_F))) ? 0 : mt.x, Un = mt.y;
    const int NT = p.NT;
    const int a_tile = kUM * kUK, b_tile = NT * kUK;                 // floats
    const int stage_floats = 2 * a_tile + 2 * b_tile;
    float* st0 = (float*)smem;
    uint64_t* bars = (uint64_t*)(smem + kUStages * (size_t)stage_floats * 4);    // [0, kUStages): stage free, [kUStages]: chunk complete
    uint32_t* s_tmem = (uint32_t*)(bars + kUStages + 1);
    const uint32_t ncols = 3 * NT <= 32 ? 32 : (3 * NT <= 64 ? 64 : (3 * NT <= 128 ? 128 : (3 * NT <= 256 ? 256 : 512)));     // MAIN | SMALL | SUM

    // rows of C that matter / whole tiles with nothing to do still have to write zeros (gradient modes)
    const int Mv = (MODE == kUmmaDG) ? ((Tn > 0) ? Un + 1 : 0) : Tn;
    const bool skip_mma = m0 >= Mv;
    if (MODE == kUmmaE && skip_mma) return;

    if (tid == 0)
        for (int i = 0; i <= kUStages; ++i) mbar_init(&bars[i], 1);
    if (warp == 0) tmem_alloc(s_tmem, ncols);
    mbar_init_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    if (!skip_mma) {
        const float* Ah = p.Ah + (size_t)n * p.a_batch + (size_t)m0 * p.lda;
        const float* Al = p.Al + (size_t)n * p.a_batch + (size_t)m0 * p.lda;
        const float* Bh = p.Bh + (size_t)n * p.b_batch + (size_t)n0 * p.ldb;
        const float* Bl = p.Bl + (size_t)n * p.b_batch + (size_t)n0 * p.ldb;
        const uint32_t idesc = umma_idesc_tf32(kUM, NT);
        const int nk = p.K / kUK;
        // element (row r, k) of a tile with R rows lives at float offset (k / 4) * R * 4 + r * 4 + (k % 4):
        // 16-byte chunk kc of row r at (kc * R + r) * 16 bytes, i.e. 8-row core matrices 128 B apart (SBO), chunks R * 16 B apart (LBO)
        auto load_tile = [&](float* dst, const float* src, int ld, int R, int k0) {
            for (int r = tid; r < R; r += 128) {
                const float* s4 = src + (size_t)r * ld + k0;
#pragma unroll
                for (int kc = 0; kc < kUK / 4; ++kc) cp_async16(dst + ((size_t)kc * R + r) * 4, s4 + 4 * kc);
            }
        };
        auto load_stage = [&](int s) {
            float* st = st0 + (size_t)(s % kUStages) * stage_floats;
            load_tile(st, Ah, p.lda, kUM, s * kUK);
            load_tile(st + a_tile, Al, p.lda, kUM, s * kUK);
            load_tile(st + 2 * a_tile, Bh, p.ldb, NT, s * kUK);
            load_tile(st + 2 * a_tile + b_tile, Bl, p.ldb, NT, s * kUK);
        };
        for (int s = 0; s < kUStages - 2; ++s) {
            if (s < nk) load_stage(s);
            cp_async_commit();
        }
        const uint32_t trow0 = tmem + ((uint32_t)(32 * warp) << 16);
        for (int s = 0; s < nk; ++s) {
            const int b = s % kUStages;
            float* st = st0 + (size_t)b * stage_floats;
            if (s + kUStages - 2 < nk) {
                // stage s + kUStages - 2 goes into the buffer the MMAs of stage s - 2 read: they were issued two
                // iterations ago, so this wait rarely blocks and the tensor core always has the next stage queued
                if (s >= 2) mbar_wait(&bars[(s - 2) % kUStages], (uint32_t)((s - 2) / kUStages) & 1u);
                load_stage(s + kUStages - 2);
            }
            cp_async_commit();
            cp_async_wait<kUStages - 2>();                                        // my copies of stage s have landed
            fence_async_smem();                                                   // generic-proxy writes -> the tensor core's reads
            tc_fence_before();
            __syncthreads();
            const bool chunk_end = ((s + 1) % kChunk == 0) || s == nk - 1;
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kUK / 8; ++ks) {
                    // one tf32 k-step = 8 elements = 2 chunks: advance the start address by 2 * LBO
                    const uint64_t ah = umma_smem_desc(st + (size_t)ks * 2 * kUM * 4, kUM * 16, 128);
                    const uint64_t al = umma_smem_desc(st + a_tile + (size_t)ks * 2 * kUM * 4, kUM * 16, 128);
                    const uint64_t bh = umma_smem_desc(st + 2 * a_tile + (size_t)ks * 2 * NT * 4, NT * 16, 128);
                    const uint64_t bl = umma_smem_desc(st + 2 * a_tile + b_tile + (size_t)ks * 2 * NT * 4, NT * 16, 128);
                    const uint32_t acc = ((s % kChunk) | ks) ? 1u : 0u;          // a chunk starts by overwriting
                    umma_tf32(tmem, ah, bh, idesc, acc);                          // MAIN
                    umma_tf32(tmem + NT, al, bh, idesc, acc);                     // SMALL
                    umma_tf32(tmem + NT, ah, bl, idesc, 1u);
                }
                umma_commit(&bars[b]);
                if (chunk_end) umma_commit(&bars[kUStages]);
            }
            if (chunk_end) {
                const int c = s / kChunk;                                         // chunk index
                mbar_wait(&bars[kUStages], (uint32_t)c & 1u);
                tc_fence_after();
                for (int c0 = 0; c0 < NT; c0 += 16) {
                    float a[16], sm[16], su[16];
                    tmem_ld16_nowait(trow0 + c0, a);
                    tmem_ld16_nowait(trow0 + NT + c0, sm);
                    if (c > 0) tmem_ld16_nowait(trow0 + 2 * NT + c0, su);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[j] = (c > 0) ? (a[j] + sm[j]) + su[j] : a[j] + sm[j];
                    tmem_st16_nowait(trow0 + 2 * NT + c0, a);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncthreads();                                                  // MAIN / SMALL may be overwritten again
                tc_fence_after();
            }
        }
    }

    // ------------------------------------------------------------------------------ epilogue ---
    const int m = m0 + 32 * warp + lane;                           // my row of C = my TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
    const int T = q.T, U1 = q.U1, V = q.V;
    if (MODE == kUmmaE) {
        const int t = m;
        const int* y = q.tgt + (size_t)n * q.Up;
        const float* f = q.f + (size_t)n * T * V;
        float2* bl = q.bl + (size_t)n * q.D * U1;
        float2* lb = q.lb + (size_t)n * q.D * U1;
        float* Eo = q.E + (size_t)n * T * U1;
        const float lf0 = (t < Tn) ? q.lf0[(size_t)n * T + t] : 0.0f, mft = (t < Tn) ? q.mf[(size_t)n * T + t] : 0.0f;
        for (int c0 = 0; c0 < NT; c0 += 16) {
            float acc[16];
            tmem_ld16(trow + 2 * NT + c0, acc);                    // SUM; every lane of the warp takes part
            if (t >= Tn) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int u = n0 + c0 + j;
                if (u > Un) continue;
                const float E = acc[j];
                Eo[(size_t)t * U1 + u] = E;
                const float lE = log2f(E);
                const size_t sk = (size_t)(t + u) * U1 + u;
                bl[sk] = log2_to_parts(lf0 + q.lg0[(size_t)n * U1 + u] - lE);
                float ll = kVoid;
                if (u < Un) ll = fmaf(f[(size_t)t * V + (y[u] & kLabelMask)], kLog2e, -mft) + q.lgy[(size_t)n * U1 + u] - lE;
                lb[sk] = log2_to_parts(ll);
            }
        }
    } else {
        const float go = q.gout[n];
        const int M = (MODE == kUmmaDF) ? T : U1;
        const bool live = !skip_mma && ((MODE == kUmmaDF) ? (m < Tn) : (m <= Un && Tn > 0));
        float* out = (MODE == kUmmaDF) ? q.gf + (size_t)n * T * V : q.gg + (size_t)n * U1 * V;
        // the factor F[m][c] (or G[m][c]) of the element-wise product is the sum of the stored halves
        const float* xh = (MODE == kUmmaDF) ? q.Fh + ((size_t)n * q.Tp + m) * V : q.Gh + ((size_t)n * q.Um + m) * V;
        const float* xl = (MODE == kUmmaDF) ? q.Fl + ((size_t)n * q.Tp + m) * V : q.Gl + ((size_t)n * q.Um + m) * V;
        // thread = accumulator row (fixed by the TMEM lane mapping), but a row-per-thread store would touch 32 rows
        // of 16 bytes per instruction: the 32 x 16 block of a warp goes through shared memory (the operand stages are
        // free by now) and leaves as 64-byte row segments, 8 rows per instruction
        float* stg = st0 + warp * (32 * 17);
        const int m_w = m0 + 32 * warp;
        (void)m; (void)xh; (void)xl; (void)live;
        for (int c0 = 0; c0 < NT; c0 += 16) {
            float acc[16];
            if (!skip_mma) {
                tmem_ld16(trow + 2 * NT + c0, acc);
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = acc[j];
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = it * 8 + (lane >> 2), cq = (lane & 3) * 4;
                const int mm = m_w + r, c = n0 + c0 + cq;
                if (mm < M && c < V) {
                    const bool lv = !skip_mma && ((MODE == kUmmaDF) ? (mm < Tn) : (mm <= Un && Tn > 0));
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lv) {
                        const size_t xo = ((MODE == kUmmaDF) ? ((size_t)n * q.Tp + mm) : ((size_t)n * q.Um + mm)) * V + c;
                        const float4 h = *(const float4*)(((MODE == kUmmaDF) ? q.Fh : q.Gh) + xo);
                        const float4 l = *(const float4*)(((MODE == kUmmaDF) ? q.Fl : q.Gl) + xo);
                        const float* a4 = stg + r * 17 + cq;
                        o = make_float4(go * a4[0] * (h.x + l.x), go * a4[1] * (h.y + l.y),
                                        go * a4[2] * (h.z + l.z), go * a4[3] * (h.w + l.w));
                    }
                    *(float4*)(out + (size_t)mm * V + c) = o;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace hab

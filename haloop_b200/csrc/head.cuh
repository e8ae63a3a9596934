// head.cuh — fused classifier head + CTC for sm_100a: replaces
//     features.log_softmax(classifier(dropout(features)))  ->  ctc loss      ha/recognizer.py:43-46, 61-82
// without ever writing the (N,T,V) logits, log-probs or their gradient to HBM.
//
// One tensor-core GEMM kernel (head_gemm_kernel) serves the four contractions of the op, told apart by its epilogue:
//   kEpiFwd    S = h W^T + b   per 128 x 128 tile: row max / sum-exp partials and the blank + label logits of the
//                              row's utterance gathered into the CTC emission rows (logits never leave the SM)
//   kEpiBwdD   S again (recomputed), d = g (softmax(S) - occupancy): written for ONE chunk of rows (an L2-sized
//                              ring, not an (N,T,V) gradient), plain and transposed
//   kEpiStore  dh = d W        plain store
//   kEpiAccum  dW += d^T h     split-K partial accumulators, one writer per element (deterministic)
//
// The GEMM: persistent CTAs, one per SM, 10 warps with fixed roles
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128 rows x 32 fp32 = one 128-byte swizzle row) of both
//               operands into a 3-stage shared-memory ring, completion on mbarriers
//   warp 1      issues tcgen05.mma kind::tf32 (M=128, N=128, K=8) from SWIZZLE_128B shared-memory descriptors into
//               TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5   split warps: kind::tf32 reads only the top 19 bits of an fp32 word, so the raw tile IS the high half;
//               these warps write the low half  x - tf32(x)  of both operand tiles (same swizzled positions, so the
//               pass is layout-blind) for the error-compensated 3-product  a b ~ ah bh + al bh + ah bl
//   warps 6-9   epilogue.  TMEM holds [MAIN0 | MAIN1 | SMALL | SUM] x 128 columns: ah bh accumulates into MAIN[c & 1] for
//               chunk c of 16 k-blocks, the two cross products (2^-11 of the magnitude) into SMALL for the whole tile; the
//               epilogue warps fold each finished chunk into SUM with round-to-nearest fp32 adds (the tensor core adds
//               into its accumulator by truncation: one accumulator over K = 1024 .. 10^4 drifts by ~half an ulp per
//               MMA), add SMALL at the last chunk, release the buffers and run the contraction's epilogue from SUM
//               while the MMA warp is already two chunks into the next tile.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace hab {

constexpr int kHM = 128, kHN = 128, kHK = 32;      // CTA tile; k elements per stage (32 fp32 = 128 B = one swizzle row)
constexpr int kHStages = 3;
constexpr int kHTile = kHM * kHK * 4;               // bytes of one operand tile (16 KB)
constexpr int kHStageBytes = 4 * kHTile;            // [A raw | B raw | A low | B low]
constexpr int kHThreads = 320;
constexpr int kHChunkKb = 16;                       // k-blocks per accumulation chunk (K = 512: 64 accumulating MMAs)
constexpr size_t kHSmem = 1024 + (size_t)kHStages * kHStageBytes + 256;

enum { kEpiFwd = 0, kEpiBwdD = 1, kEpiStore = 2, kEpiAccum = 3 };
constexpr int kHasDup = 0x40000000;           // cls2pos flag: the class occurs again later in the target

struct HeadGemmParams {
    int M, N, K;                 // rows of A this launch covers, rows of B (= output columns), contraction length
    int a_row0;                  // added to A's row coordinate (start of the row chunk)
    int tiles_m, tiles_n, splits, kb_per_split;
    int nprod;                   // 3: error-compensated tf32 x 3; 1: plain tf32
    // epilogue data
    const float* bias;           // (N) or null
    int rows_total, T;           // all rows of h (= Nutt * T); frames per utterance
    const int4* meta; const int* cls2pos; const int* dupnext; int Sp, V;
    float* em; int E;            // emission rows (fwd: raw logits gathered in; bwd: occupancy rows)
    float2* stats;               // fwd: [tiles_n][rows_total] (max, sum exp) partials
    const float* lse2; const float* loss; const float* gout;
    float* out; long long ldo;   // kEpiBwdD: d (M x N chunk), kEpiStore: dh, kEpiAccum: partials [splits][M][N]
    float* outT; long long ldt;  // kEpiBwdD: d^T (N x ldt)
    int accumulate;              // kEpiAccum: add to what is there (every chunk but the first)
};

// K-major SWIZZLE_128B shared-memory descriptor: rows 128 B apart, 8-row groups 1024 B apart (SBO); the k-step inside the
// 128-byte row is selected by advancing the start address (the hardware applies the XOR swizzle to the address bits)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(kHThreads, 1)
head_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, HeadGemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;                   // SWIZZLE_128B tiles: 1024-byte aligned
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = (uint64_t*)(sm + (size_t)kHStages * kHStageBytes);
    uint64_t* raw_full = bars;                      // [kHStages] TMA -> split warps, MMA
    uint64_t* lo_full = bars + kHStages;            // [kHStages] split warps -> MMA
    uint64_t* empty = bars + 2 * kHStages;          // [kHStages] MMA -> TMA
    uint64_t* acc_full = bars + 3 * kHStages;       // [2] MMA -> epilogue: chunk finished in MAIN[b]
    uint64_t* acc_empty = acc_full + 2;             // [2] epilogue -> MMA: MAIN[b] folded into SUM
    uint64_t* small_empty = acc_empty + 2;          // [1] epilogue -> MMA: SMALL folded into SUM (once per tile)
    uint32_t* s_tmem = (uint32_t*)(small_empty + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kHStages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&lo_full[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        mbar_init(small_empty, 128);
    }
    if (warp == 1) tmem_alloc(s_tmem, 512);
    mbar_init_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const int ntiles = p.tiles_m * p.tiles_n * p.splits;
    const int nkb_all = (p.K + kHK - 1) / kHK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int nb = tile % p.tiles_n, mb = (tile / p.tiles_n) % p.tiles_m, sp = tile / (p.tiles_n * p.tiles_m);
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kHStages;
                    mbar_wait(&empty[s], ((it / kHStages) & 1u) ^ 1u);
                    mbar_expect_tx(&raw_full[s], 2u * kHTile);
                    const uint32_t st = base + (uint32_t)s * kHStageBytes;
                    tma_load_2d(st, &mapA, kb * kHK, p.a_row0 + mb * kHM, &raw_full[s]);
                    tma_load_2d(st + kHTile, &mapB, kb * kHK, nb * kHN, &raw_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kHM, kHN);
            uint32_t it = 0, lt = 0, gc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                const int sp = tile / (p.tiles_n * p.tiles_m);
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                const uint32_t d_small = tmem + 256u;
                for (int kc = kb0; kc < kb1; kc += kHChunkKb, ++gc) {
                    // a chunk of kHChunkKb k-blocks accumulates into MAIN[gc & 1]; the epilogue warps fold it into SUM
                    const uint32_t b = gc & 1u;
                    mbar_wait(&acc_empty[b], ((gc >> 1) & 1u) ^ 1u);
                    if (kc == kb0 && p.nprod == 3) mbar_wait(small_empty, (lt & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_main = tmem + b * 128u;
                    const int kce = min(kb1, kc + kHChunkKb);
                    for (int kb = kc; kb < kce; ++kb, ++it) {
                        const int s = it % kHStages;
                        const uint32_t ph = (it / kHStages) & 1u;
                        mbar_wait(&raw_full[s], ph);
                        if (p.nprod == 3) mbar_wait(&lo_full[s], ph);
                        tc_fence_after();
                        const uint32_t st = base + (uint32_t)s * kHStageBytes;
#pragma unroll
                        for (int ks = 0; ks < kHK / 8; ++ks) {
                            const uint64_t ah = umma_desc_sw128(st + ks * 32), bh = umma_desc_sw128(st + kHTile + ks * 32);
                            umma_tf32(d_main, ah, bh, idesc, (kb > kc || ks > 0) ? 1u : 0u);
                            if (p.nprod == 3) {
                                const uint64_t al = umma_desc_sw128(st + 2 * kHTile + ks * 32), bl = umma_desc_sw128(st + 3 * kHTile + ks * 32);
                                umma_tf32(d_small, al, bh, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                                umma_tf32(d_small, ah, bl, idesc, 1u);
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                    umma_commit(&acc_full[b]);
                }
            }
        }
    } else if (warp < 6) {
        // low halves: 2048 16-byte chunks of [A raw | B raw] -> [A low | B low], 16 per thread
        const int ct = tid - 64;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int sp = tile / (p.tiles_n * p.tiles_m);
            const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % kHStages;
                if (p.nprod != 3) continue;
                mbar_wait(&raw_full[s], (it / kHStages) & 1u);
                const float4* src = (const float4*)(sm + (size_t)s * kHStageBytes);
                float4* dst = (float4*)(sm + (size_t)s * kHStageBytes + 2 * kHTile);
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const float4 v = src[ct + 128 * i];
                    dst[ct + 128 * i] = make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
                }
                fence_async_smem();
                mbar_arrive(&lo_full[s]);
            }
        }
    } else {
        const int q = warp & 3;                                  // TMEM lanes [32 q, 32 q + 32) belong to this warp
        uint32_t lt = 0, gc = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int nb = tile % p.tiles_n, mb = (tile / p.tiles_n) % p.tiles_m, sp = tile / (p.tiles_n * p.tiles_m);
            const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
            const uint32_t trow = tlane + 384u;                   // SUM: what the epilogue below reads
            {
                // fold every finished chunk into SUM with round-to-nearest fp32 adds (the tensor core accumulates by
                // truncation: one accumulator over a long K drifts by ~half an ulp per MMA); SMALL joins at the last one
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                for (int kc = kb0; kc < kb1; kc += kHChunkKb, ++gc) {
                    const uint32_t b = gc & 1u;
                    const bool first = kc == kb0, last = kc + kHChunkKb >= kb1;
                    if (lane == 0) mbar_wait(&acc_full[b], (gc >> 1) & 1u);
                    __syncwarp();
                    tc_fence_after();
                    for (int c0 = 0; c0 < kHN && nb * kHN + c0 < p.N; c0 += 16) {
                        float a[16], su[16], sm[16];
                        tmem_ld16_nowait(tlane + b * 128u + c0, a);
                        if (!first) tmem_ld16_nowait(trow + c0, su);
                        if (last && p.nprod == 3) tmem_ld16_nowait(tlane + 256u + c0, sm);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float r = a[j];
                            if (last && p.nprod == 3) r += sm[j];
                            if (!first) r += su[j];
                            a[j] = r;
                        }
                        tmem_st16_nowait(trow + c0, a);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&acc_empty[b]);
                    if (last && p.nprod == 3) mbar_arrive(small_empty);
                }
            }
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
                        if (rl < p.M) {
                            float* drow = p.out + (size_t)rl * p.ldo + n0 + c0;
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4)
                                if (n0 + c0 + 4 * j4 < p.N)
                                    *(float4*)(drow + 4 * j4) = make_float4(dv[4 * j4], dv[4 * j4 + 1], dv[4 * j4 + 2], dv[4 * j4 + 3]);
                        }
                        // transposed copy: the 32 lanes of a warp are 32 consecutive rows -> one 128-byte line per column
                        if (rl < p.ldt) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (n0 + c0 + j < p.N) p.outT[(size_t)(n0 + c0 + j) * p.ldt + rl] = dv[j];
                        }
                    }
                }
                if (EPI == kEpiFwd && inside) p.stats[(size_t)nb * p.rows_total + row] = make_float2(m, ssum);
            } else {
                float* orow = (EPI == kEpiAccum) ? p.out + ((size_t)sp * p.M + rl) * p.ldo + n0
                                                 : p.out + ((size_t)p.a_row0 + rl) * p.ldo + n0;
                for (int c0 = 0; c0 < kHN && n0 + c0 < p.N; c0 += 16) {
                    float a[16];
                    tmem_ld16_nowait(trow + c0, a);
                    tmem_ld_wait();
                    if (rl < p.M && (EPI == kEpiAccum || p.a_row0 + rl < p.rows_total)) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            if (n0 + c0 + 4 * j4 >= p.N) continue;
                            float4 o;
                            o.x = a[4 * j4]; o.y = a[4 * j4 + 1]; o.z = a[4 * j4 + 2]; o.w = a[4 * j4 + 3];
                            float4* dst = (float4*)(orow + c0 + 4 * j4);
                            if (EPI == kEpiAccum && p.accumulate) {
                                const float4 old = *dst;
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *dst = o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

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

// grid (ceil(C / 32), ceil(Rd / 32)), block (32, 8): dst[c][r] = src[r0 + r][c] for r0 + r < Rs (else 0), r < Rd
__global__ void __launch_bounds__(256) head_transpose_kernel(const float* src, long long r0, long long Rs, int C, float* dst, int Rd) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, rb = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const long long r = r0 + rb + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < Rs && c < C) ? src[(size_t)r * C + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = rb + threadIdx.x;
        if (c < C && r < Rd) dst[(size_t)c * Rd + r] = tile[threadIdx.x][i];
    }
}

// grid ceil(V / 8), block 256 (warp per class): db[c] (+)= sum over the chunk's rows of d^T[c][:]
__global__ void __launch_bounds__(256) head_colsum_kernel(const float* dT, int ldt, int rows, int V, float* db, int accumulate) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= V) return;
    const float* r = dT + (size_t)c * ldt;
    float s = 0.0f;
    for (int i = lane; i < rows; i += 32) s += r[i];
    s = warp_sum(s);
    if (lane == 0) db[c] = accumulate ? db[c] + s : s;
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

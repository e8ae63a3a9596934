// umma_gemm.cuh — the tensor-core GEMM engine shared by the fused classifier head (head.cuh) and the joint-free RNN-T
// contractions (rnnt_fg_umma.cuh):  C (M x N) = A (M x K) B^T (N x K), both operands K-major fp32 in global memory,
// products on tcgen05 kind::tf32, error-compensated by a three-product split, fp32 results handed to an epilogue functor.
//
// Persistent CTAs, one per SM, 10 warps with fixed roles
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D tiles (128 rows x 32 fp32 = one 128-byte swizzle row) of both
//               operands into a 3-stage shared-memory ring, completion on mbarriers
//   warp 1      issues tcgen05.mma kind::tf32 (M=128, N=128, K=8) from SWIZZLE_128B shared-memory descriptors into
//               TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-5   split warps: kind::tf32 reads only the top 19 bits of an fp32 word, so the raw tile IS the high half;
//               these warps write the low half  x - tf32(x)  of both operand tiles (same swizzled positions, so the
//               pass is layout-blind) for the error-compensated 3-product  a b ~ ah bh + al bh + ah bl
//   warps 6-9   epilogue.  TMEM holds [MAIN0 | MAIN1 | SMALL | SUM] x 128 columns: ah bh accumulates into MAIN[c & 1] for
//               chunk c of `chunk_kb` k-blocks, the two cross products (2^-11 of the magnitude) into SMALL for the whole
//               tile; the epilogue warps fold each finished chunk into SUM with round-to-nearest fp32 adds (the tensor
//               core adds into its accumulator by truncation: one accumulator over K = 1024 .. 10^4 drifts by ~half an
//               ulp per MMA), add SMALL at the last chunk, release the buffers and run the caller's epilogue from SUM
//               while the MMA warp is already two chunks into the next tile.
// Tiles are (batch, split-K part, row block, column block); batches are independent problems at row offsets of the same
// two tensor maps.  Template kernel + inline device functions only: safe to include from several translation units.
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
constexpr size_t kHSmem = 1024 + (size_t)kHStages * kHStageBytes + 256 + 4 * 32 * 17 * 4;

// K-major SWIZZLE_128B shared-memory descriptor: rows 128 B apart, 8-row groups 1024 B apart (SBO); the k-step inside the
// 128-byte row is selected by advancing the start address (the hardware applies the XOR swizzle to the address bits)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// MN-major descriptor for a 128 (MN) x 32 (K) fp32 tile stored as four 32 (MN) x 32 (K) boxes of 4096 bytes, each written
// by TMA with the 32-byte-atom 128-byte swizzle (layout type SWIZZLE_128B_BASE32B = 1): a K row is one 128-byte line of
// 32 consecutive MN elements, 4 K rows form a 512-byte swizzle atom.  LBO = distance between the atoms along MN (the
// boxes: 4096 B), SBO = distance between the 4-row atoms along K (512 B).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (1ull << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// What the GEMM engine itself needs to know (the epilogue functor carries the rest).
struct GemmCore {
    int N, K;                    // output columns (rows of B), contraction length
    int a_row0;                  // added to A's row coordinate
    int tiles_m, tiles_n, splits, kb_per_split;
    int batches, a_batch_rows, b_batch_rows;     // independent problems: row offsets of batch b in the A and B tensor maps
    // An operand may also be given "MN-major": stored (K x MN) row-major, i.e. the matrix the caller already has when the
    // contraction runs over its ROW index (d^T h, W^T F): no transposed copy is made, the tile is loaded as four
    // 32 (MN) x 32 (K) boxes and the MMA reads it through an MN-major descriptor.  For such an operand the tensor map has
    // box 32 x 32 and the 32-byte-atom 128-byte swizzle (the only MN-major layout kind::tf32 accepts); `*_row0` /
    // `*_batch_rows` offset the MN coordinate (columns) and `*_k0` / `*_batch_k` the K coordinate (rows).
    int a_mn, b_mn;
    int a_k0, b_k0, a_batch_k, b_batch_k, b_row0;
    int nprod;                   // 3: error-compensated tf32 x 3; 1: plain tf32
    int chunk_kb;                // k-blocks per accumulation chunk
};

template <class Epi>
__global__ void __launch_bounds__(kHThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmCore p, const Epi epi) {
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
    float* stage = (float*)(sm + (size_t)kHStages * kHStageBytes + 256);     // 4 x (32 x 17) floats: epilogue staging, one block per warp
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
    const int ntiles = p.tiles_m * p.tiles_n * p.splits * p.batches;
    const int nkb_all = (p.K + kHK - 1) / kHK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int nb = tile % p.tiles_n, mb = (tile / p.tiles_n) % p.tiles_m, z = tile / (p.tiles_n * p.tiles_m), sp = z % p.splits, bt = z / p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kHStages;
                    mbar_wait(&empty[s], ((it / kHStages) & 1u) ^ 1u);
                    mbar_expect_tx(&raw_full[s], 2u * kHTile);
                    const uint32_t st = base + (uint32_t)s * kHStageBytes;
                    const int ka = p.a_k0 + bt * p.a_batch_k + kb * kHK, ra = p.a_row0 + bt * p.a_batch_rows + mb * kHM;
                    const int kbb = p.b_k0 + bt * p.b_batch_k + kb * kHK, rb = p.b_row0 + bt * p.b_batch_rows + nb * kHN;
                    if (!p.a_mn) tma_load_2d(st, &mapA, ka, ra, &raw_full[s]);
                    else
#pragma unroll
                        for (int i = 0; i < 4; ++i) tma_load_2d(st + i * 4096, &mapA, ra + 32 * i, ka, &raw_full[s]);
                    if (!p.b_mn) tma_load_2d(st + kHTile, &mapB, kbb, rb, &raw_full[s]);
                    else
#pragma unroll
                        for (int i = 0; i < 4; ++i) tma_load_2d(st + kHTile + i * 4096, &mapB, rb + 32 * i, kbb, &raw_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kHM, kHN) | (p.a_mn ? (1u << 15) : 0u) | (p.b_mn ? (1u << 16) : 0u);
            uint32_t it = 0, lt = 0, gc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                const int sp = (tile / (p.tiles_n * p.tiles_m)) % p.splits;
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                const uint32_t d_small = tmem + 256u;
                for (int kc = kb0; kc < kb1; kc += p.chunk_kb, ++gc) {
                    // a chunk of p.chunk_kb k-blocks accumulates into MAIN[gc & 1]; the epilogue warps fold it into SUM
                    const uint32_t b = gc & 1u;
                    mbar_wait(&acc_empty[b], ((gc >> 1) & 1u) ^ 1u);
                    if (kc == kb0 && p.nprod == 3) mbar_wait(small_empty, (lt & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_main = tmem + b * 128u;
                    const int kce = min(kb1, kc + p.chunk_kb);
                    for (int kb = kc; kb < kce; ++kb, ++it) {
                        const int s = it % kHStages;
                        const uint32_t ph = (it / kHStages) & 1u;
                        mbar_wait(&raw_full[s], ph);
                        if (p.nprod == 3) mbar_wait(&lo_full[s], ph);
                        tc_fence_after();
                        const uint32_t st = base + (uint32_t)s * kHStageBytes;
#pragma unroll
                        for (int ks = 0; ks < kHK / 8; ++ks) {
                            // k-step inside the stage: 32 bytes along the 128-byte row (K-major) or 8 rows of 128 bytes (MN-major)
                            const uint32_t oa = p.a_mn ? ks * 1024 : ks * 32, ob = p.b_mn ? ks * 1024 : ks * 32;
                            const uint64_t ah = p.a_mn ? umma_desc_sw128_mn(st + oa) : umma_desc_sw128(st + oa);
                            const uint64_t bh = p.b_mn ? umma_desc_sw128_mn(st + kHTile + ob) : umma_desc_sw128(st + kHTile + ob);
                            umma_tf32(d_main, ah, bh, idesc, (kb > kc || ks > 0) ? 1u : 0u);
                            if (p.nprod == 3) {
                                const uint64_t al = p.a_mn ? umma_desc_sw128_mn(st + 2 * kHTile + oa) : umma_desc_sw128(st + 2 * kHTile + oa);
                                const uint64_t bl = p.b_mn ? umma_desc_sw128_mn(st + 3 * kHTile + ob) : umma_desc_sw128(st + 3 * kHTile + ob);
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
            const int sp = (tile / (p.tiles_n * p.tiles_m)) % p.splits;
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
            const int nb = tile % p.tiles_n, mb = (tile / p.tiles_n) % p.tiles_m, z = tile / (p.tiles_n * p.tiles_m), sp = z % p.splits, bt = z / p.splits;
            const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
            const uint32_t trow = tlane + 384u;                   // SUM: what the epilogue below reads
            {
                // fold every finished chunk into SUM with round-to-nearest fp32 adds (the tensor core accumulates by
                // truncation: one accumulator over a long K drifts by ~half an ulp per MMA); SMALL joins at the last one
                const int kb0 = sp * p.kb_per_split, kb1 = min(nkb_all, kb0 + p.kb_per_split);
                for (int kc = kb0; kc < kb1; kc += p.chunk_kb, ++gc) {
                    const uint32_t b = gc & 1u;
                    const bool first = kc == kb0, last = kc + p.chunk_kb >= kb1;
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
            epi(trow, mb, nb, sp, bt, q, lane, stage + (warp - 6) * (32 * 17));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace hab

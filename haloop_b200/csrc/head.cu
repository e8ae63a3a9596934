// head.cu — host side of the fused classifier head + CTC (head.cuh): workspace layout, tensor maps, launches.
// C ABI: ha_head_ctc_workspace_bytes / ha_head_ctc_fwd / ha_head_ctc_bwd (include/ha_b200.h).
#include <cuda.h>

#include "../../include/ha_b200.h"
#include "head.cuh"
#include "host.h"

namespace hab {

namespace {

struct HeadWs {
    // saved (forward -> backward)
    size_t meta, order, tgt, dupnext, cls2pos, loss, lse2, em, saved_total;
    // forward scratch
    size_t stats, tr, fwd_total;
    // backward scratch
    size_t d, part, colpart, bwd_total;
    int Sp, E, JWp, SPX, tiles_n, R, splits_max;
};

int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return n;
}

HeadWs head_ws_layout(int N, int T, int D, int V, int S) {
    HeadWs w{};
    ctc_head_dims(S, &w.Sp, &w.E, &w.JWp, &w.SPX);
    const size_t rows = (size_t)N * T;
    w.tiles_n = (V + kHN - 1) / kHN;
    size_t o = 256;
    auto take = [&](size_t bytes) { size_t at = o; o = round_up_sz(o + bytes, 256); return at; };
    w.meta = take(sizeof(int4) * (size_t)N);
    w.order = take(sizeof(int) * (size_t)N);
    w.tgt = take(sizeof(int) * (size_t)N * w.Sp);
    w.dupnext = take(sizeof(int) * (size_t)N * w.Sp);
    w.cls2pos = take(sizeof(int) * (size_t)N * V);
    w.loss = take(sizeof(float) * (size_t)N);
    w.lse2 = take(sizeof(float) * rows);
    w.em = take(sizeof(float) * rows * w.E);
    w.saved_total = o;
    o = 256;
    w.stats = take(sizeof(float2) * rows * w.tiles_n);
    w.tr = take(sizeof(float) * rows * w.SPX);
    w.fwd_total = o;
    // backward: rows are processed in chunks of R so that d of a chunk (and the h rows it pairs with) stay in L2 (~40 MB).
    // A chunk's GEMMs are separate launches of persistent CTAs, so R is chosen to make their tile counts whole
    // multiples of the SM count (a 512-tile launch on 148 SMs idles 13 % of the machine in its last wave).
    const int nsm = sm_count();
    const size_t widest = (size_t)(V > D ? V : D);
    long long blocks = (long long)((40u << 20) / (4 * widest)) / kHM;          // row blocks of 128 that fit the budget
    if (blocks < 1) blocks = 1;
    {
        auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
        const int unit = nsm / gcd(nsm, w.tiles_n);                             // row blocks per whole wave pattern
        if (blocks >= unit) blocks = blocks / unit * unit;
    }
    long long R = blocks * kHM;
    const long long rows_up = (long long)round_up_sz(rows, kHM);
    if (R > rows_up) R = rows_up;
    w.R = (int)R;
    // split-K of dW = d^T h over the chunk's rows: the split count that wastes the least of the last wave, each split
    // keeping at least 16 k-blocks
    const int tiles_dw = ((V + kHM - 1) / kHM) * ((D + kHN - 1) / kHN);
    const int nkb = (int)(R / kHK);
    int best = 1;
    double best_eff = 0.0;
    for (int sp = 1; sp <= 16 && (sp == 1 || nkb / sp >= 16); ++sp) {
        const long long t = (long long)tiles_dw * sp;
        const double eff = (double)t / (double)(((t + nsm - 1) / nsm) * nsm);
        if (eff > best_eff + 0.02) { best_eff = eff; best = sp; }
    }
    w.splits_max = best;
    o = 256;
    const int Vp = round_up(V, 4);                                   // leading dimension of the class-contiguous scratch matrices
    w.d = take(sizeof(float) * (size_t)w.R * Vp);
    w.part = take(sizeof(float) * (size_t)w.splits_max * V * D);
    w.colpart = take(sizeof(float) * (size_t)kColSplits * V);
    w.bwd_total = o;
    return w;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}
// (rows x cols) fp32 matrix, cols contiguous (the contraction index), leading dimension ld floats; box 128 rows x 32 cols
int make_map(CUtensorMap* m, const float* ptr, size_t rows, size_t cols, size_t ld, int mn = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return host_fail(HA_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kHK, (cuuint32_t)(mn ? 32 : kHM)};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return host_fail(HA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%zu cols=%zu ld=%zu", (int)r, rows, cols, ld);
    return HA_OK;
}

template <int EPI>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, HeadGemmParams p, cudaStream_t st, const char* what) {
    static thread_local bool done[8] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 7]) {
        cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<HeadEpi<EPI>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmem);
        if (e != cudaSuccess) return host_fail(HA_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
        done[dev & 7] = true;
    }
    GemmCore c{};
    c.N = p.N; c.K = p.K; c.a_row0 = p.a_row0; c.nprod = p.nprod; c.chunk_kb = kHChunkKb;
    c.batches = 1; c.a_batch_rows = 0; c.b_batch_rows = 0;
    c.a_mn = p.a_mn; c.b_mn = p.b_mn; c.a_k0 = p.a_k0; c.b_k0 = p.b_k0; c.a_batch_k = 0; c.b_batch_k = 0; c.b_row0 = 0;
    c.tiles_m = (p.M + kHM - 1) / kHM;
    c.tiles_n = (p.N + kHN - 1) / kHN;
    const int nkb = (p.K + kHK - 1) / kHK;
    c.splits = p.splits < 1 ? 1 : (p.splits > nkb ? nkb : p.splits);
    c.kb_per_split = (nkb + c.splits - 1) / c.splits;
    c.splits = (nkb + c.kb_per_split - 1) / c.kb_per_split;          // no empty split
    const int ntiles = c.tiles_m * c.tiles_n * c.splits;
    const int grid = ntiles < sm_count() ? ntiles : sm_count();
    HeadEpi<EPI> epi{p};
    umma_gemm_kernel<HeadEpi<EPI>><<<grid, kHThreads, kHSmem, st>>>(a, b, c, epi);
    return host_check_launch(what);
}

int check_shapes(int N, int T, int D, int V, int S) {
    if (N <= 0 || T <= 0 || D <= 0 || V <= 0 || S < 0) return host_fail(HA_ERR_INVALID_ARGUMENT, "bad sizes N=%d T=%d D=%d V=%d S=%d", N, T, D, V, S);
    if (N > 65535) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "N=%d > 65535", N);
    if (D % 4) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "the feature dimension must be a multiple of 4 (D=%d): rows are copied by TMA", D);
    if ((long long)N * T > 0x7fffff00ll) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "N*T too large");
    if (S > 1023) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "target length > 1023 is not supported");
    return HA_OK;
}

}  // namespace

int host_make_map(void* map, const float* ptr, size_t rows, size_t cols, size_t ld, int mn) { return make_map((CUtensorMap*)map, ptr, rows, cols, ld, mn); }
int host_sm_count() { return sm_count(); }

}  // namespace hab

using namespace hab;

extern "C" {

int ha_head_ctc_workspace_bytes(int N, int T, int D, int V, int S, size_t* saved, size_t* fwd_scratch, size_t* bwd_scratch) {
    int rc = check_shapes(N, T, D, V, S);
    if (rc) return rc;
    const HeadWs w = head_ws_layout(N, T, D, V, S);
    if (saved) *saved = w.saved_total;
    if (fwd_scratch) *fwd_scratch = w.fwd_total;
    if (bwd_scratch) *bwd_scratch = w.bwd_total;
    return HA_OK;
}

int ha_head_ctc_fwd(const float* h, const float* W, const float* bias, int N, int T, int D, int V,
                    const void* targets, int64_t tgt_stride, int S, int targets_i64,
                    const void* in_len, const void* tgt_len, int lengths_i64, int precision,
                    float* loss, void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, void* stream) {
    int rc = check_shapes(N, T, D, V, S);
    if (rc) return rc;
    if (!h || !W || !in_len || !tgt_len || !loss || !saved || !scratch || (S > 0 && !targets)) return host_fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (!aligned16(h) || !aligned16(W) || !aligned16(saved) || !aligned16(scratch)) return host_fail(HA_ERR_INVALID_ARGUMENT, "buffers must be 16-byte aligned");
    if (precision != 1 && precision != 3) return host_fail(HA_ERR_INVALID_ARGUMENT, "precision must be 1 (tf32) or 3 (tf32 x 3)");
    const HeadWs w = head_ws_layout(N, T, D, V, S);
    if (saved_bytes < w.saved_total || scratch_bytes < w.fwd_total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* sb = (unsigned char*)saved;
    unsigned char* cb = (unsigned char*)scratch;
    const int rows = N * T;

    if ((rc = ctc_prep_for_head(targets, tgt_stride, S, targets_i64, in_len, tgt_len, lengths_i64, T, N, V, w.Sp,
                                sb + w.meta, (int*)(sb + w.order), (int*)(sb + w.tgt), (int*)(sb + w.dupnext), st))) return rc;
    head_cls2pos_kernel<<<N, 256, 0, st>>>((const int4*)(sb + w.meta), (const int*)(sb + w.tgt), (const int*)(sb + w.dupnext), w.Sp, V, (int*)(sb + w.cls2pos));
    if ((rc = host_check_launch("head_cls2pos_kernel"))) return rc;

    CUtensorMap ma, mb;
    if ((rc = make_map(&ma, h, (size_t)rows, (size_t)D, (size_t)D))) return rc;
    if ((rc = make_map(&mb, W, (size_t)V, (size_t)D, (size_t)D))) return rc;
    HeadGemmParams p{};
    p.M = rows; p.N = V; p.K = D; p.a_row0 = 0; p.splits = 1; p.nprod = precision;
    p.bias = bias; p.rows_total = rows; p.T = T;
    p.meta = (const int4*)(sb + w.meta); p.cls2pos = (const int*)(sb + w.cls2pos); p.dupnext = (const int*)(sb + w.dupnext);
    p.Sp = w.Sp; p.V = V; p.em = (float*)(sb + w.em); p.E = w.E; p.stats = (float2*)(cb + w.stats);
    if ((rc = launch_gemm<kEpiFwd>(ma, mb, p, st, "head_gemm_kernel<fwd>"))) return rc;

    HeadFinalizeParams fp{};
    fp.rows_total = rows; fp.T = T; fp.tiles_n = w.tiles_n; fp.meta = p.meta; fp.stats = p.stats;
    fp.lse2 = (float*)(sb + w.lse2); fp.em = p.em; fp.E = w.E;
    head_finalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(fp);
    if ((rc = host_check_launch("head_finalize_kernel"))) return rc;

    return ctc_trellis_for_head(T, N, S, w.Sp, w.E, w.SPX, w.JWp, sb + w.meta, (const int*)(sb + w.order), (const int*)(sb + w.tgt),
                                p.em, (float*)(cb + w.tr), loss, (float*)(sb + w.loss), st);
}

int ha_head_ctc_bwd(const float* h, const float* W, const float* bias, int N, int T, int D, int V, int S,
                    const float* grad_loss, int precision, float* dh, float* dW, float* db,
                    void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, void* stream) {
    int rc = check_shapes(N, T, D, V, S);
    if (rc) return rc;
    if (!h || !W || !grad_loss || !dh || !dW || !saved || !scratch) return host_fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (!aligned16(h) || !aligned16(W) || !aligned16(dh) || !aligned16(dW) || !aligned16(saved) || !aligned16(scratch))
        return host_fail(HA_ERR_INVALID_ARGUMENT, "buffers must be 16-byte aligned");
    if (precision != 1 && precision != 3) return host_fail(HA_ERR_INVALID_ARGUMENT, "precision must be 1 (tf32) or 3 (tf32 x 3)");
    const HeadWs w = head_ws_layout(N, T, D, V, S);
    if (saved_bytes < w.saved_total || scratch_bytes < w.bwd_total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* sb = (unsigned char*)saved;
    unsigned char* cb = (unsigned char*)scratch;
    const int rows = N * T;
    float* d = (float*)(cb + w.d);
    float* part = (float*)(cb + w.part);
    float* colpart = (float*)(cb + w.colpart);
    const int R = w.R, Vp = round_up(V, 4);

    CUtensorMap m_h, m_W, m_d, m_W_mn, m_d_mn, m_h_mn;
    if ((rc = make_map(&m_h, h, (size_t)rows, (size_t)D, (size_t)D))) return rc;
    if ((rc = make_map(&m_W, W, (size_t)V, (size_t)D, (size_t)D))) return rc;
    if ((rc = make_map(&m_d, d, (size_t)R, (size_t)V, (size_t)Vp))) return rc;
    // dh = d W contracts over the ROWS of W (classes): W is read as it lies (MN-major B operand)
    if ((rc = make_map(&m_W_mn, W, (size_t)V, (size_t)D, (size_t)D, 1))) return rc;
    // dW = d^T h contracts over the ROWS of d and h: both are read as they lie (MN-major operands), no transposed copies
    if ((rc = make_map(&m_d_mn, d, (size_t)R, (size_t)V, (size_t)Vp, 1))) return rc;
    if ((rc = make_map(&m_h_mn, h, (size_t)rows, (size_t)D, (size_t)D, 1))) return rc;

    int splits_used = 1;
    for (int r0 = 0, chunk = 0; r0 < rows; r0 += R, ++chunk) {
        const int cr = (rows - r0 < R) ? rows - r0 : R;              // rows of this chunk
        const int crp = round_up(cr, kHM);
        HeadGemmParams p{};
        p.M = cr; p.N = V; p.K = D; p.a_row0 = r0; p.splits = 1; p.nprod = precision;
        p.bias = bias; p.rows_total = rows; p.T = T;
        p.meta = (const int4*)(sb + w.meta); p.cls2pos = (const int*)(sb + w.cls2pos); p.dupnext = (const int*)(sb + w.dupnext);
        p.Sp = w.Sp; p.V = V; p.em = (float*)(sb + w.em); p.E = w.E;
        p.lse2 = (const float*)(sb + w.lse2); p.loss = (const float*)(sb + w.loss); p.gout = grad_loss;
        p.out = d; p.ldo = Vp; p.Mpad = crp;                         // rows cr .. crp of d are written as zeros (dW reads them)
        if ((rc = launch_gemm<kEpiBwdD>(m_h, m_W, p, st, "umma_gemm_kernel<head bwd d>"))) return rc;
        if (db) {
            head_colsum_kernel<<<dim3((V + 31) / 32, kColSplits), dim3(32, 8), 0, st>>>(d, Vp, cr, V, colpart);
            if ((rc = host_check_launch("head_colsum_kernel"))) return rc;
            head_colsum_finish_kernel<<<(V + 255) / 256, 256, 0, st>>>(colpart, V, db, chunk > 0);
            if ((rc = host_check_launch("head_colsum_finish_kernel"))) return rc;
        }
        HeadGemmParams q{};
        q.M = cr; q.N = D; q.K = V; q.a_row0 = 0; q.splits = 1; q.nprod = precision; q.rows_total = cr;
        q.out = dh + (size_t)r0 * D; q.ldo = D; q.b_mn = 1;
        if ((rc = launch_gemm<kEpiStore>(m_d, m_W_mn, q, st, "umma_gemm_kernel<head dh>"))) return rc;
        HeadGemmParams u{};
        u.M = V; u.N = D; u.K = crp; u.a_row0 = 0; u.nprod = precision; u.rows_total = V;
        u.a_mn = 1; u.b_mn = 1; u.a_k0 = 0; u.b_k0 = r0;
        u.splits = w.splits_max; u.out = part; u.ldo = D; u.accumulate = chunk > 0;
        // a shorter last chunk cuts K differently; each partial buffer still has exactly one writer per launch
        if ((rc = launch_gemm<kEpiAccum>(m_d_mn, m_h_mn, u, st, "umma_gemm_kernel<head dW>"))) return rc;
        if (chunk == 0) {
            const int nkb = (crp + kHK - 1) / kHK;
            int s = w.splits_max > nkb ? nkb : w.splits_max;
            const int kbps = (nkb + s - 1) / s;
            splits_used = (nkb + kbps - 1) / kbps;
        }
    }
    const size_t nW = (size_t)V * D;
    head_reduce_kernel<<<(unsigned)((nW + 255) / 256), 256, 0, st>>>(part, splits_used, nW, dW);
    return host_check_launch("head_reduce_kernel");
}

}  // extern "C"

// api.cu — the C ABI of libha_b200.so (include/ha_b200.h): argument checks, workspace carving,
// kernel selection and launches.  No allocation, no synchronisation, no global state.
#include <cuda.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_set>

#include "../../include/ha_b200.h"
#include "align.cuh"
#include "beam.cuh"
#include "common.cuh"
#include "ctc.cuh"
#include "host.h"
#include "rnnt.cuh"
#include "rnnt_fg.cuh"
#include "rnnt_fg_umma.cuh"
#include "star.cuh"

using namespace hab;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HA_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return HA_OK;
}

constexpr size_t kMaxSmem = hab::kMaxSmemOptin;
constexpr size_t kRowSmemTarget = 100 * 1024; // two row-kernel CTAs per SM

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is sticky per (kernel, device): it is raised
// to the opt-in maximum the first time this thread launches the kernel on a device and not touched again.
template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
    if (bytes > kMaxSmem) return fail(HA_ERR_UNSUPPORTED_SHAPE, "%s needs %zu B of shared memory", what, bytes);
    thread_local std::unordered_set<unsigned long long> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long key = (unsigned long long)(uintptr_t)(const void*)kernel * 64ull + (unsigned)(dev & 63);
    if (done.count(key)) return HA_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return fail(HA_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    done.insert(key);
    return HA_OK;
}

using hab::aligned16;

// warps per CTA and ring depth for a warp-per-row streaming kernel with `row_floats` per stage
struct RowCfg { int nwarps, nstage, rows_per_warp; size_t smem; bool ok; };

template <typename F>
RowCfg pick_row_cfg(int T, F smem_of /* (nstage, nwarps) -> bytes */, int ns_max = 4, int nw_max = 8) {
    // (ns_max, nw_max) = (2, 4) for the statistics ("rows") kernels and the CTC / star gradient kernels: a warp's
    // per-row dependency chain (max, sum, log, gather / scatter) is long, so many small CTAs per SM (5-6 of four
    // warps with a two-stage TMA ring each) hide more latency than a deep ring under few warps.  Measured on
    // B200 (nw, ns sweep): RNN-T rows 1.20 -> 0.97 ms, CTC rows 0.44 -> 0.37 ms, CTC grad 0.62 -> 0.56 ms, star
    // grad 0.22 -> 0.14 ms; the RNN-T gradient kernel (already at the copy peak) keeps 8 warps x 3 stages.
    RowCfg c{8, 4, 1, 0, false};
    for (int nw = nw_max; nw >= 1 && !c.ok; nw >>= 1) {
        for (int ns = ns_max; ns >= 2; --ns) {
            size_t b = smem_of(ns, nw);
            if (b <= kRowSmemTarget || (ns == 2 && b <= kMaxSmem)) { c.nwarps = nw; c.nstage = ns; c.smem = b; c.ok = true; break; }
        }
    }
    int rpw = (T + c.nwarps * 4 - 1) / (c.nwarps * 4);
    int cap = 16;
    c.rows_per_warp = rpw < 1 ? 1 : (rpw > cap ? cap : rpw);
    return c;
}

int common_checks(const void* x, int T, int N, int V, int S, const void* ws, size_t ws_bytes, size_t need) {
    if (!x || !ws) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (T <= 0 || N <= 0 || V <= 0 || S < 0) return fail(HA_ERR_INVALID_ARGUMENT, "bad sizes T=%d N=%d V=%d S=%d", T, N, V, S);
    if (N > 65535) return fail(HA_ERR_UNSUPPORTED_SHAPE, "N=%d > 65535", N);
    if (!aligned16(ws)) return fail(HA_ERR_INVALID_ARGUMENT, "workspace must be 16-byte aligned");
    if (ws_bytes < need) return fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, need);
    return HA_OK;
}

// one contraction of the joint-free RNN-T on the GEMM engine: every utterance is a batch entry of the same two tensor maps.
// An operand is (rows x cols) per utterance, row-major; K-major (mn = 0): rows index the output, cols the contraction;
// MN-major (mn = 1): rows index the contraction, cols the output (the matrix is used as it lies, no transposed copy).
struct FgOperand { const float* ptr; size_t rows, cols; int mn; };
template <int MODE>
int fg_engine_launch(FgOperand A, FgOperand B, int batches, int tiles_m, int ncols, int K, const FgUmmaParams& q,
                     cudaStream_t st, const char* what) {
    alignas(64) CUtensorMap ma, mb;
    int rc;
    if ((rc = host_make_map(&ma, A.ptr, (size_t)batches * A.rows, A.cols, A.cols, A.mn))) return rc;
    if ((rc = host_make_map(&mb, B.ptr, (size_t)batches * B.rows, B.cols, B.cols, B.mn))) return rc;
    if ((rc = set_smem(umma_gemm_kernel<FgEpi<MODE>>, kHSmem, what))) return rc;
    GemmCore c{};
    c.N = ncols; c.K = K; c.a_row0 = 0; c.nprod = 3; c.chunk_kb = kFgChunkKb;
    c.tiles_m = tiles_m; c.tiles_n = (ncols + kHN - 1) / kHN; c.splits = 1; c.kb_per_split = (K + kHK - 1) / kHK;
    c.batches = batches;
    c.a_mn = A.mn; c.b_mn = B.mn;
    c.a_batch_rows = A.mn ? 0 : (int)A.rows; c.a_batch_k = A.mn ? (int)A.rows : 0;      // the batch offsets the ROW coordinate
    c.b_batch_rows = B.mn ? 0 : (int)B.rows; c.b_batch_k = B.mn ? (int)B.rows : 0;
    const int ntiles = c.tiles_m * c.tiles_n * batches;
    const int grid = ntiles < host_sm_count() ? ntiles : host_sm_count();
    FgEpi<MODE> epi{q};
    umma_gemm_kernel<FgEpi<MODE>><<<grid, kHThreads, kHSmem, st>>>(ma, mb, c, epi);
    return check_launch(what);
}

}  // namespace

namespace hab {
int host_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int host_check_launch(const char* what) { return check_launch(what); }
}  // namespace hab

extern "C" {

int ha_b200_version(void) { return 100; }
const char* ha_b200_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------- CTC ---
// Two implementations share the entry points: the fused two-kernel path (ctc2.cu; V a multiple of 4 and rows that
// fit the shared-memory rings) and the three-kernel path below (any V, unaligned views).  The choice depends on
// (T, N, V, S) only, so the workspace query, the forward and the backward call agree on the layout.
size_t ha_ctc_workspace_bytes(int T, int N, int V, int S) {
    if (T <= 0 || N <= 0 || S < 0) return 0;
    if (ctc2_eligible(T, N, V, S)) return ctc2_workspace_bytes(T, N, S);
    return ctc_ws_layout(T, N, S).total;
}

static int ctc_trellis_launch(const TrellisParams& tp, int nslot, int N, cudaStream_t st) {
    TrellisParams p = tp;
    // One CTA per utterance: W compute warps per sweep direction, each owning J slots of 32 label
    // pairs, plus one producer warp per direction.  (J, W) are template parameters (the per-step
    // barriers and mailbox indices become immediates); two slots per warp up to 5 warps measured best.
    if (nslot > 32) return fail(HA_ERR_UNSUPPORTED_SHAPE, "target length > 1023 is not supported");
    int W = nslot < 2 ? 1 : ((nslot + 1) / 2 < 5 ? (nslot + 1) / 2 : 5);
    int J = (nslot + W - 1) / W;
    if (J == 7) J = 8;
    if (J > 8) return fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: J=%d", J);
    W = (nslot + J - 1) / J;
    p.W = W;
    p.G = kMaxG;
    int ns = 4;
    while (ns >= 2 && (size_t)2 * trellis_dir_bytes(p.E, p.SPX, 4 + p.Sp, ns, p.G, p.W, 32) > 100 * 1024) --ns;
    if (ns < 2) {
        ns = 2;
        while (p.G > 1 && (size_t)2 * trellis_dir_bytes(p.E, p.SPX, 4 + p.Sp, ns, p.G, p.W, 32) > 218 * 1024) p.G >>= 1;
    }
    p.nstage = ns;
    p.dir_bytes = trellis_dir_bytes(p.E, p.SPX, 4 + p.Sp, ns, p.G, p.W, 32);
    const size_t smem = (size_t)2 * p.dir_bytes + kTrellisGuard;
    const dim3 grid(N), block(32 * (2 * p.W + 2));
    int rc = HA_ERR_UNSUPPORTED_SHAPE;
    bool hit = false;
#define HAB_TRY(JJ, WW)                                                                        \
    if (!hit && J == JJ && W == WW) {                                                          \
        hit = true;                                                                            \
        if ((rc = set_smem(ctc_trellis_kernel<JJ, WW>, smem, "ctc_trellis"))) return rc;       \
        ctc_trellis_kernel<JJ, WW><<<grid, block, smem, st>>>(p);                              \
    }
#define HAB_TRY_J(JJ) HAB_TRY(JJ, 1) HAB_TRY(JJ, 2) HAB_TRY(JJ, 3) HAB_TRY(JJ, 4) HAB_TRY(JJ, 5)
    HAB_TRY_J(1) HAB_TRY_J(2) HAB_TRY_J(3) HAB_TRY_J(4) HAB_TRY_J(5) HAB_TRY_J(6) HAB_TRY_J(8)
#undef HAB_TRY_J
#undef HAB_TRY
    if (!hit) return fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: J=%d W=%d", J, W);
    return check_launch("ctc_trellis_kernel");
}

}  // extern "C"

namespace hab {
void ctc_head_dims(int S, int* Sp, int* E, int* JWp, int* SPX) {
    const CtcWs w = ctc_ws_layout(1, 1, S);
    *Sp = w.Sp; *E = w.E; *JWp = w.JWp; *SPX = w.SPX;
}
int ctc_prep_for_head(const void* targets, int64_t tgt_stride, int S, int targets_i64,
                      const void* in_len, const void* tgt_len, int lengths_i64, int T, int N, int V, int Sp,
                      void* meta, int* order, int* tgt, int* dupnext, cudaStream_t st) {
    PrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = Sp;
    pp.meta = (int4*)meta; pp.order = order; pp.tgt = tgt; pp.dupnext = dupnext; pp.star = 0;
    ctc_prep_kernel<<<N, 256, (size_t)round_up(S > 0 ? S : 1, 4) * 4, st>>>(pp);
    return check_launch("ctc_prep_kernel");
}
int ctc_trellis_for_head(int T, int N, int S, int Sp, int E, int SPX, int JWp, const void* meta, const int* order,
                         const int* tgt, float* em, float* tr, float* loss, float* loss_ws, cudaStream_t st) {
    TrellisParams tp{};
    tp.T = T; tp.N = N; tp.meta = (const int4*)meta; tp.order = order; tp.tgt = tgt; tp.Sp = Sp;
    tp.em = em; tp.E = E; tp.tr = tr; tp.SPX = SPX; tp.JWp = JWp; tp.loss = loss; tp.loss_ws = loss_ws;
    return ctc_trellis_launch(tp, (S + 1 + 31) / 32, N, st);
}
}  // namespace hab

extern "C" {

int ha_ctc_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
               const void* targets, int64_t tgt_stride, int S, int targets_i64,
               const void* in_len, const void* tgt_len, int lengths_i64,
               int from_logits, float* loss, void* ws, size_t ws_bytes, void* stream) {
    if (ctc2_eligible(T, N, V, S)) {
        int rc2 = common_checks(x, T, N, V, S, ws, ws_bytes, 0);
        if (rc2) return rc2;
        if (!in_len || !tgt_len || !loss || (S > 0 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
        return ctc2_fwd(x, sx_t, sx_n, T, N, V, targets, tgt_stride, S, targets_i64, in_len, tgt_len, lengths_i64,
                        from_logits, loss, ws, ws_bytes, (cudaStream_t)stream);
    }
    const CtcWs w = ctc_ws_layout(T, N, S);
    int rc = common_checks(x, T, N, V, S, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!in_len || !tgt_len || !loss || (S > 0 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;

    PrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = w.Sp;
    pp.meta = (int4*)(base + w.meta); pp.order = (int*)(base + w.order);
    pp.tgt = (int*)(base + w.tgt); pp.dupnext = (int*)(base + w.dupnext); pp.star = 0;
    ctc_prep_kernel<<<N, 256, (size_t)round_up(S > 0 ? S : 1, 4) * 4, st>>>(pp);
    if ((rc = check_launch("ctc_prep_kernel"))) return rc;

    RowsParams rp{};
    rp.x = x; rp.sx_t = sx_t; rp.sx_n = sx_n; rp.T = T; rp.N = N; rp.V = V;
    rp.meta = pp.meta; rp.tgt = pp.tgt; rp.Sp = w.Sp;
    rp.lse2 = (float*)(base + w.lse2); rp.em = (float*)(base + w.em); rp.E = w.E;
    rp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(x) && (sx_t % 4 == 0) && (sx_n % 4 == 0);
    rp.use_bulk = vec ? 1 : 0;
    RowCfg rc_ = pick_row_cfg(T, [&](int ns, int nw) { return rows_smem_bytes(w.Sp, V, ns, nw); }, 2, 4);
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the row kernel", V);
    rp.nstage = rc_.nstage; rp.nwarps = rc_.nwarps; rp.rows_per_warp = rc_.rows_per_warp;
    {
        const dim3 grid((T + rp.nwarps * rp.rows_per_warp - 1) / (rp.nwarps * rp.rows_per_warp), N);
        const dim3 block(32 * rp.nwarps);
        if (vec) {
            if ((rc = set_smem(ctc_rows_kernel<true>, rc_.smem, "ctc_rows"))) return rc;
            ctc_rows_kernel<true><<<grid, block, rc_.smem, st>>>(rp);
        } else {
            if ((rc = set_smem(ctc_rows_kernel<false>, rc_.smem, "ctc_rows"))) return rc;
            ctc_rows_kernel<false><<<grid, block, rc_.smem, st>>>(rp);
        }
        if ((rc = check_launch("ctc_rows_kernel"))) return rc;
    }

    TrellisParams tp{};
    tp.T = T; tp.N = N; tp.meta = pp.meta; tp.order = pp.order; tp.tgt = pp.tgt; tp.Sp = w.Sp;
    tp.em = rp.em; tp.E = w.E; tp.tr = (float*)(base + w.tr); tp.SPX = w.SPX; tp.JWp = w.JWp;
    tp.loss = loss; tp.loss_ws = (float*)(base + w.loss);
    if ((rc = ctc_trellis_launch(tp, (S + 1 + 31) / 32, N, st))) return rc;

    return HA_OK;
}

int ha_ctc_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
               const float* grad_loss, int from_logits,
               float* gx, int64_t sg_t, int64_t sg_n,
               void* ws, size_t ws_bytes, void* stream) {
    if (ctc2_eligible(T, N, V, S)) {
        int rc2 = common_checks(gx, T, N, V, S, ws, ws_bytes, 0);
        if (rc2) return rc2;
        if (!grad_loss) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
        return ctc2_bwd(x, sx_t, sx_n, T, N, V, S, grad_loss, from_logits, gx, sg_t, sg_n, ws, ws_bytes, (cudaStream_t)stream);
    }
    const CtcWs w = ctc_ws_layout(T, N, S);
    int rc = common_checks(gx, T, N, V, S, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!grad_loss || (from_logits && !x)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;
    GradParams gp{};
    gp.x = x; gp.sx_t = sx_t; gp.sx_n = sx_n; gp.gx = gx; gp.sg_t = sg_t; gp.sg_n = sg_n;
    gp.T = T; gp.N = N; gp.V = V;
    gp.meta = (const int4*)(base + w.meta); gp.tgt = (const int*)(base + w.tgt);
    gp.dupnext = (const int*)(base + w.dupnext); gp.Sp = w.Sp;
    gp.lse2 = (const float*)(base + w.lse2); gp.occ = (const float*)(base + w.em); gp.E = w.E;
    gp.gout = grad_loss; gp.loss = (const float*)(base + w.loss); gp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(gx) && (sg_t % 4 == 0) && (sg_n % 4 == 0) &&
                     (!from_logits || (aligned16(x) && (sx_t % 4 == 0) && (sx_n % 4 == 0)));
    gp.use_bulk = vec ? 1 : 0;
    RowCfg rc_ = pick_row_cfg(T, [&](int ns, int nw) { return grad_smem_bytes(w.Sp, V, w.E, ns, nw); }, 2, 4);
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the gradient kernel", V);
    gp.nstage = rc_.nstage; gp.nwarps = rc_.nwarps; gp.rows_per_warp = rc_.rows_per_warp;
    const dim3 grid((T + gp.nwarps * gp.rows_per_warp - 1) / (gp.nwarps * gp.rows_per_warp), N);
    const dim3 block(32 * gp.nwarps);
    if (vec) {
        if ((rc = set_smem(ctc_grad_kernel<true>, rc_.smem, "ctc_grad"))) return rc;
        ctc_grad_kernel<true><<<grid, block, rc_.smem, st>>>(gp);
    } else {
        if ((rc = set_smem(ctc_grad_kernel<false>, rc_.smem, "ctc_grad"))) return rc;
        ctc_grad_kernel<false><<<grid, block, rc_.smem, st>>>(gp);
    }
    return check_launch("ctc_grad_kernel");
}

// -------------------------------------------------------------------------------- star-CTC ---
size_t ha_star_workspace_bytes(int T, int N, int V, int S) {
    if (T <= 0 || N <= 0 || S < 0) return 0;
    if (star2_eligible(T, N, V, S)) return star2_workspace_bytes(T, N, S);
    return star_ws_layout(T, N, S).total;
}

static int star_trellis_launch(const StarTrellisParams& tp, int nslot, int N, cudaStream_t st) {
    StarTrellisParams p = tp;
    if (nslot > 16) return fail(HA_ERR_UNSUPPORTED_SHAPE, "star-CTC target length > 511 is not supported");
    int W = nslot < 2 ? 1 : (nslot < 4 ? 2 : 4);
    const int J = (nslot + W - 1) / W;
    if (J > 4) return fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: star J=%d", J);
    W = (nslot + J - 1) / J;
    p.W = W;
    p.G = kMaxG;
    const int OC = 4 + 2 * p.Sp;
    int ns = 4;
    while (ns >= 2 && (size_t)2 * trellis_dir_bytes(p.E, p.SPX, OC, ns, p.G, p.W, 64) > 100 * 1024) --ns;
    if (ns < 2) {
        ns = 2;
        while (p.G > 1 && (size_t)2 * trellis_dir_bytes(p.E, p.SPX, OC, ns, p.G, p.W, 64) > 220 * 1024) p.G >>= 1;
    }
    p.nstage = ns;
    p.dir_bytes = trellis_dir_bytes(p.E, p.SPX, OC, ns, p.G, p.W, 64);
    const size_t smem = (size_t)2 * p.dir_bytes;
    const dim3 grid(N), block(32 * (2 * p.W + 2));
    int rc = HA_ERR_UNSUPPORTED_SHAPE;
    bool hit = false;
#define HAB_TRY(JJ, WW)                                                                         \
    if (!hit && J == JJ && W == WW) {                                                           \
        hit = true;                                                                             \
        if ((rc = set_smem(star_trellis_kernel<JJ, WW>, smem, "star_trellis"))) return rc;      \
        star_trellis_kernel<JJ, WW><<<grid, block, smem, st>>>(p);                              \
    }
#define HAB_TRY_J(JJ) HAB_TRY(JJ, 1) HAB_TRY(JJ, 2) HAB_TRY(JJ, 3) HAB_TRY(JJ, 4) HAB_TRY(JJ, 5)
    HAB_TRY_J(1) HAB_TRY_J(2) HAB_TRY_J(3) HAB_TRY_J(4)
#undef HAB_TRY_J
#undef HAB_TRY
    if (!hit) return fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: star J=%d W=%d", J, W);
    return check_launch("star_trellis_kernel");
}

int ha_star_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
                const void* targets, int64_t tgt_stride, int S, int targets_i64,
                const void* in_len, const void* tgt_len, int lengths_i64,
                float star_penalty, int from_logits, float* loss,
                void* ws, size_t ws_bytes, void* stream) {
    if (star2_eligible(T, N, V, S)) {
        int rc2 = common_checks(x, T, N, V, S, ws, ws_bytes, 0);
        if (rc2) return rc2;
        if (!in_len || !tgt_len || !loss || (S > 0 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
        return star2_fwd(x, sx_t, sx_n, T, N, V, targets, tgt_stride, S, targets_i64, in_len, tgt_len, lengths_i64,
                         star_penalty, from_logits, loss, ws, ws_bytes, (cudaStream_t)stream);
    }
    const StarWs w = star_ws_layout(T, N, S);
    int rc = common_checks(x, T, N, V, S, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!in_len || !tgt_len || !loss || (S > 0 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (V < 2) return fail(HA_ERR_UNSUPPORTED_SHAPE, "star-CTC needs at least one non-blank class");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;

    PrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = w.Sp;
    pp.meta = (int4*)(base + w.meta); pp.order = (int*)(base + w.order);
    pp.tgt = (int*)(base + w.tgt); pp.dupnext = (int*)(base + w.dupnext); pp.star = 1;
    ctc_prep_kernel<<<N, 256, (size_t)round_up(S > 0 ? S : 1, 4) * 4, st>>>(pp);
    if ((rc = check_launch("ctc_prep_kernel"))) return rc;

    StarRowsParams rp{};
    rp.x = x; rp.sx_t = sx_t; rp.sx_n = sx_n; rp.T = T; rp.N = N; rp.V = V; rp.S = S;
    rp.meta = pp.meta; rp.tgt = pp.tgt; rp.Sp = w.Sp;
    rp.lse2 = (float*)(base + w.lse2); rp.em = (float*)(base + w.em); rp.E = w.E;
    rp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(x) && (sx_t % 4 == 0) && (sx_n % 4 == 0);
    rp.use_bulk = vec ? 1 : 0;
    RowCfg rc_ = pick_row_cfg(T, [&](int ns, int nw) { return rows_smem_bytes(w.Sp, V, ns, nw); }, 2, 4);
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the row kernel", V);
    rp.nstage = rc_.nstage; rp.nwarps = rc_.nwarps; rp.rows_per_warp = rc_.rows_per_warp;
    {
        const dim3 grid((T + rp.nwarps * rp.rows_per_warp - 1) / (rp.nwarps * rp.rows_per_warp), N);
        const dim3 block(32 * rp.nwarps);
        if (vec) {
            if ((rc = set_smem(star_rows_kernel<true>, rc_.smem, "star_rows"))) return rc;
            star_rows_kernel<true><<<grid, block, rc_.smem, st>>>(rp);
        } else {
            if ((rc = set_smem(star_rows_kernel<false>, rc_.smem, "star_rows"))) return rc;
            star_rows_kernel<false><<<grid, block, rc_.smem, st>>>(rp);
        }
        if ((rc = check_launch("star_rows_kernel"))) return rc;
    }

    StarTrellisParams tp{};
    tp.T = T; tp.N = N; tp.S = S; tp.meta = pp.meta; tp.order = pp.order; tp.tgt = pp.tgt; tp.Sp = w.Sp;
    tp.em = rp.em; tp.E = w.E; tp.tr = (float*)(base + w.tr); tp.SPX = w.SPX; tp.JWp = w.JWp;
    tp.loss = loss; tp.loss_ws = (float*)(base + w.loss);
    tp.pen2 = star_penalty * kLog2e;
    return star_trellis_launch(tp, (S + 1 + 31) / 32, N, st);
}

int ha_star_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
                const float* grad_loss, int from_logits,
                float* gx, int64_t sg_t, int64_t sg_n,
                void* ws, size_t ws_bytes, void* stream) {
    if (star2_eligible(T, N, V, S)) {
        int rc2 = common_checks(gx, T, N, V, S, ws, ws_bytes, 0);
        if (rc2) return rc2;
        if (!grad_loss) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
        return star2_bwd(x, sx_t, sx_n, T, N, V, S, grad_loss, from_logits, gx, sg_t, sg_n, ws, ws_bytes, (cudaStream_t)stream);
    }
    const StarWs w = star_ws_layout(T, N, S);
    int rc = common_checks(gx, T, N, V, S, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!grad_loss || !x) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;
    StarGradParams gp{};
    gp.x = x; gp.sx_t = sx_t; gp.sx_n = sx_n; gp.gx = gx; gp.sg_t = sg_t; gp.sg_n = sg_n;
    gp.T = T; gp.N = N; gp.V = V; gp.S = S;
    gp.meta = (const int4*)(base + w.meta); gp.tgt = (const int*)(base + w.tgt);
    gp.dupnext = (const int*)(base + w.dupnext); gp.Sp = w.Sp;
    gp.lse2 = (const float*)(base + w.lse2); gp.occ = (const float*)(base + w.em); gp.E = w.E;
    gp.gout = grad_loss; gp.loss = (const float*)(base + w.loss); gp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(gx) && (sg_t % 4 == 0) && (sg_n % 4 == 0) &&
                     aligned16(x) && (sx_t % 4 == 0) && (sx_n % 4 == 0);
    gp.use_bulk = vec ? 1 : 0;
    RowCfg rc_ = pick_row_cfg(T, [&](int ns, int nw) { return grad_smem_bytes(w.Sp, V, w.E, ns, nw); }, 2, 4);
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the gradient kernel", V);
    gp.nstage = rc_.nstage; gp.nwarps = rc_.nwarps; gp.rows_per_warp = rc_.rows_per_warp;
    const dim3 grid((T + gp.nwarps * gp.rows_per_warp - 1) / (gp.nwarps * gp.rows_per_warp), N);
    const dim3 block(32 * gp.nwarps);
    if (vec) {
        if ((rc = set_smem(star_grad_kernel<true>, rc_.smem, "star_grad"))) return rc;
        star_grad_kernel<true><<<grid, block, rc_.smem, st>>>(gp);
    } else {
        if ((rc = set_smem(star_grad_kernel<false>, rc_.smem, "star_grad"))) return rc;
        star_grad_kernel<false><<<grid, block, rc_.smem, st>>>(gp);
    }
    return check_launch("star_grad_kernel");
}

// ----------------------------------------------------------------------------------- RNN-T ---
size_t ha_rnnt_workspace_bytes(int N, int T, int U1, int V) {
    (void)V;
    if (T <= 0 || N <= 0 || U1 <= 0) return 0;
    return rnnt_ws_layout(N, T, U1).total;
}

int ha_rnnt_fwd(const float* joint, int64_t sj_n, int64_t sj_t, int64_t sj_u, int N, int T, int U1, int V,
                const void* targets, int64_t tgt_stride, int targets_i64,
                const void* in_len, const void* tgt_len, int lengths_i64,
                int from_logits, float* loss, void* ws, size_t ws_bytes, void* stream) {
    if (U1 <= 0) return fail(HA_ERR_INVALID_ARGUMENT, "U1 must be >= 1");
    const RnntWs w = rnnt_ws_layout(N, T, U1);
    int rc = common_checks(joint, T, N, V, U1 - 1, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!in_len || !tgt_len || !loss || (U1 > 1 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (U1 > 1024) return fail(HA_ERR_UNSUPPORTED_SHAPE, "U+1 > 1024 is not supported");
    if ((long long)T * U1 >= (1ll << 31)) return fail(HA_ERR_UNSUPPORTED_SHAPE, "T*(U+1) too large");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;

    RnntPrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.N = N; pp.T = T; pp.U = U1 - 1; pp.V = V; pp.Up = w.Up;
    pp.meta = (int4*)(base + w.meta); pp.tgt = (int*)(base + w.tgt);
    rnnt_prep_kernel<<<N, 128, 0, st>>>(pp);
    if ((rc = check_launch("rnnt_prep_kernel"))) return rc;

    RnntRowsParams rp{};
    rp.x = joint; rp.sx_n = sj_n; rp.sx_t = sj_t; rp.sx_u = sj_u; rp.N = N; rp.T = T; rp.U1 = U1; rp.V = V;
    rp.meta = pp.meta; rp.tgt = pp.tgt; rp.Up = w.Up;
    rp.lse2 = (float*)(base + w.lse2); rp.bl = (float2*)(base + w.bl); rp.lb = (float2*)(base + w.lb); rp.D = w.D;
    rp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(joint) && (sj_n % 4 == 0) && (sj_t % 4 == 0) && (sj_u % 4 == 0);
    rp.use_bulk = vec ? 1 : 0;
    const int nodes = T * U1;
    RowCfg rc_ = pick_row_cfg(nodes, [&](int ns, int nw) { return rnnt_rows_smem_bytes(V, ns, nw); }, 2, 4);
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the row kernel", V);
    rp.nstage = rc_.nstage; rp.nwarps = rc_.nwarps; rp.rows_per_warp = rc_.rows_per_warp;
    {
        const dim3 grid((nodes + rp.nwarps * rp.rows_per_warp - 1) / (rp.nwarps * rp.rows_per_warp), N);
        const dim3 block(32 * rp.nwarps);
        if (vec) {
            if ((rc = set_smem(rnnt_rows_kernel<true>, rc_.smem, "rnnt_rows"))) return rc;
            rnnt_rows_kernel<true><<<grid, block, rc_.smem, st>>>(rp);
        } else {
            if ((rc = set_smem(rnnt_rows_kernel<false>, rc_.smem, "rnnt_rows"))) return rc;
            rnnt_rows_kernel<false><<<grid, block, rc_.smem, st>>>(rp);
        }
        if ((rc = check_launch("rnnt_rows_kernel"))) return rc;
    }

    RnntLatticeParams lp{};
    lp.N = N; lp.T = T; lp.U1 = U1; lp.D = w.D; lp.meta = pp.meta;
    lp.bl = rp.bl; lp.lb = rp.lb; lp.alpha = (double*)(base + w.alpha); lp.beta = (double*)(base + w.beta); lp.occ = (float2*)(base + w.occ);
    lp.loss = loss; lp.loss_ws = (float*)(base + w.loss);
    {
        const int half = round_up(U1, 32), nthreads = half * (half <= 512 ? 2 : 1);
        if (nthreads <= 256) rnnt_lattice_kernel<256><<<N, nthreads, 0, st>>>(lp);
        else rnnt_lattice_kernel<1024><<<N, nthreads, 0, st>>>(lp);
    }
    return check_launch("rnnt_lattice_kernel");
}

int ha_rnnt_bwd(const float* joint, int64_t sj_n, int64_t sj_t, int64_t sj_u, int N, int T, int U1, int V,
                const float* grad_loss, int from_logits, float* gjoint, int64_t sg_n, int64_t sg_t, int64_t sg_u,
                void* ws, size_t ws_bytes, void* stream) {
    if (U1 <= 0) return fail(HA_ERR_INVALID_ARGUMENT, "U1 must be >= 1");
    const RnntWs w = rnnt_ws_layout(N, T, U1);
    int rc = common_checks(gjoint, T, N, V, U1 - 1, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!grad_loss || (from_logits && !joint)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;
    RnntGradParams gp{};
    gp.x = joint; gp.gx = gjoint; gp.N = N; gp.T = T; gp.U1 = U1; gp.V = V;
    gp.sx_n = sj_n; gp.sx_t = sj_t; gp.sx_u = sj_u; gp.sg_n = sg_n; gp.sg_t = sg_t; gp.sg_u = sg_u;
    gp.meta = (const int4*)(base + w.meta); gp.tgt = (const int*)(base + w.tgt); gp.Up = w.Up;
    gp.lse2 = (const float*)(base + w.lse2); gp.occ = (const float2*)(base + w.occ); gp.D = w.D;
    gp.gout = grad_loss; gp.loss = (const float*)(base + w.loss); gp.from_logits = from_logits;
    const bool vec = (V % 4 == 0) && aligned16(gjoint) && (sg_n % 4 == 0) && (sg_t % 4 == 0) && (sg_u % 4 == 0) &&
                     (!from_logits || (aligned16(joint) && (sj_n % 4 == 0) && (sj_t % 4 == 0) && (sj_u % 4 == 0)));
    gp.use_bulk = vec ? 1 : 0;
    const int nodes = T * U1;
    RowCfg rc_ = pick_row_cfg(nodes, [&](int ns, int nw) { return rnnt_rows_smem_bytes(V, ns, nw); });
    if (!rc_.ok) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V=%d too large for the gradient kernel", V);
    gp.nstage = rc_.nstage; gp.nwarps = rc_.nwarps; gp.rows_per_warp = rc_.rows_per_warp;
    {
        const dim3 grid((nodes + 63) / 64, N);
        rnnt_zero_kernel<<<grid, 256, 0, st>>>(gp);
        if ((rc = check_launch("rnnt_zero_kernel"))) return rc;
    }
    const dim3 grid((nodes + gp.nwarps * gp.rows_per_warp - 1) / (gp.nwarps * gp.rows_per_warp), N);
    const dim3 block(32 * gp.nwarps);
    if (vec) {
        if ((rc = set_smem(rnnt_grad_kernel<true>, rc_.smem, "rnnt_grad"))) return rc;
        rnnt_grad_kernel<true><<<grid, block, rc_.smem, st>>>(gp);
    } else {
        if ((rc = set_smem(rnnt_grad_kernel<false>, rc_.smem, "rnnt_grad"))) return rc;
        rnnt_grad_kernel<false><<<grid, block, rc_.smem, st>>>(gp);
    }
    return check_launch("rnnt_grad_kernel");
}

// ------------------------------------------------------------- joint-free (factored) RNN-T ---
// The three contractions run on the tensor cores (rnnt_fg_umma.cuh) when the class count tiles evenly; the fp32 SIMT
// kernels of rnnt_fg.cuh remain for every other shape.
static int fg_umma_nt(int V) {            // tensor-core path of the joint-free RNN-T: rows of F / G must be TMA-copyable
    return (V % 16 == 0) ? 128 : 0;       // (16-byte aligned rows; a multiple of 16 keeps every padded extent aligned too)
}

size_t ha_rnnt_fg_workspace_bytes(int N, int T, int U1, int V) {
    if (T <= 0 || N <= 0 || U1 <= 0) return 0;
    size_t total = rnnt_fg_ws_layout(N, T, U1).total;
    if (fg_umma_nt(V)) total += fg_umma_ws_layout(N, T, U1, V).total;
    return total;
}

static FgUmmaParams fg_umma_params(const RnntFgParams& r, const FgUmmaWs& u, unsigned char* ubase) {
    FgUmmaParams q{};
    q.f = r.f; q.g = r.g; q.gf = r.gf; q.gg = r.gg;
    q.N = r.N; q.T = r.T; q.U1 = r.U1; q.V = r.V; q.Up = r.Up; q.D = r.D;
    q.meta = r.meta; q.tgt = r.tgt;
    q.mf = r.mf; q.mg = r.mg; q.lf0 = r.lf0; q.lg0 = r.lg0; q.lgy = r.lgy; q.E = r.E;
    q.bl = r.bl; q.lb = r.lb; q.occ = r.occ; q.gout = r.gout; q.loss = r.loss;
    q.Fh = (float*)(ubase + u.Fh); q.Fl = (float*)(ubase + u.Fl); q.Gh = (float*)(ubase + u.Gh); q.Gl = (float*)(ubase + u.Gl);
    q.Fth = (float*)(ubase + u.Fth); q.Ftl = (float*)(ubase + u.Ftl); q.Gth = (float*)(ubase + u.Gth); q.Gtl = (float*)(ubase + u.Gtl);
    q.Wh = (float*)(ubase + u.Wh); q.Wl = (float*)(ubase + u.Wl); q.Wth = (float*)(ubase + u.Wth); q.Wtl = (float*)(ubase + u.Wtl);
    q.Tp = u.Tp; q.Tk = u.Tk; q.Uk = u.Uk; q.Um = u.Um;
    return q;
}

static RnntFgParams rnnt_fg_params(const float* f, const float* g, int N, int T, int U1, int V,
                                   const RnntFgWs& w, unsigned char* base) {
    RnntFgParams p{};
    p.f = f; p.g = g; p.N = N; p.T = T; p.U1 = U1; p.V = V; p.Up = w.Up; p.D = w.D;
    p.meta = (const int4*)(base + w.meta); p.tgt = (int*)(base + w.tgt); p.nxt = (int*)(base + w.nxt);
    p.mf = (float*)(base + w.mf); p.mg = (float*)(base + w.mg);
    p.lf0 = (float*)(base + w.lf0); p.lg0 = (float*)(base + w.lg0); p.lgy = (float*)(base + w.lgy);
    p.E = (float*)(base + w.E);
    p.bl = (float2*)(base + w.bl); p.lb = (float2*)(base + w.lb); p.occ = (const float2*)(base + w.occ);
    p.loss = (const float*)(base + w.loss);
    return p;
}

int ha_rnnt_fg_fwd(const float* f, const float* g, int N, int T, int U1, int V,
                   const void* targets, int64_t tgt_stride, int targets_i64,
                   const void* in_len, const void* tgt_len, int lengths_i64,
                   float* loss, void* ws, size_t ws_bytes, void* stream) {
    if (U1 <= 0) return fail(HA_ERR_INVALID_ARGUMENT, "U1 must be >= 1");
    const RnntFgWs w = rnnt_fg_ws_layout(N, T, U1);
    int rc = common_checks(f, T, N, V, U1 - 1, ws, ws_bytes, ha_rnnt_fg_workspace_bytes(N, T, U1, V));
    if (rc) return rc;
    if (!g || !in_len || !tgt_len || !loss || (U1 > 1 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (U1 > 1024) return fail(HA_ERR_UNSUPPORTED_SHAPE, "U+1 > 1024 is not supported");
    if ((long long)(T + U1) * U1 >= (1ll << 31)) return fail(HA_ERR_UNSUPPORTED_SHAPE, "T*(U+1) too large");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;

    RnntPrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.N = N; pp.T = T; pp.U = U1 - 1; pp.V = V; pp.Up = w.Up;
    pp.meta = (int4*)(base + w.meta); pp.tgt = (int*)(base + w.tgt);
    rnnt_prep_kernel<<<N, 128, 0, st>>>(pp);
    if ((rc = check_launch("rnnt_prep_kernel"))) return rc;

    RnntFgParams p = rnnt_fg_params(f, g, N, T, U1, V, w, base);
    rnnt_fg_chain_kernel<<<N, 128, (size_t)w.Up * 4, st>>>(p);
    if ((rc = check_launch("rnnt_fg_chain_kernel"))) return rc;
    if (fg_umma_nt(V) && aligned16(f) && aligned16(g)) {
        const FgUmmaWs uw = fg_umma_ws_layout(N, T, U1, V);
        const FgUmmaParams q = fg_umma_params(p, uw, base + w.total);
        fg_rows_kernel<<<dim3((uw.Tp + uw.Um + 7) / 8, N), 256, 0, st>>>(q);
        if ((rc = check_launch("fg_rows_kernel"))) return rc;
        // E = F G^T: rows of F (Tp per utterance) x rows of G (Um per utterance, Uk columns of E used), K = V
        if ((rc = fg_engine_launch<kUmmaE>({q.Fh, (size_t)uw.Tp, (size_t)V, 0}, {q.Gh, (size_t)uw.Um, (size_t)V, 0}, N, uw.Tp / kHM, uw.Uk, V, q, st,
                                           "umma_gemm_kernel<fg E>"))) return rc;
        fg_arc_kernel<<<dim3((T + 7) / 8, N), 256, 0, st>>>(q);
        if ((rc = check_launch("fg_arc_kernel"))) return rc;
    } else {
        rnnt_fg_stats_kernel<<<dim3((T + U1 + 7) / 8, N), 256, 0, st>>>(p);
        if ((rc = check_launch("rnnt_fg_stats_kernel"))) return rc;
        rnnt_fg_gemm_kernel<kE><<<dim3((T + kGM - 1) / kGM, (U1 + kGN - 1) / kGN, N), 256, 0, st>>>(p);
        if ((rc = check_launch("rnnt_fg_gemm_kernel<E>"))) return rc;
    }

    RnntLatticeParams lp{};
    lp.N = N; lp.T = T; lp.U1 = U1; lp.D = w.D; lp.meta = p.meta;
    lp.bl = p.bl; lp.lb = p.lb; lp.alpha = (double*)(base + w.alpha); lp.beta = (double*)(base + w.beta);
    lp.occ = (float2*)(base + w.occ);
    lp.loss = loss; lp.loss_ws = (float*)(base + w.loss);
    {
        const int half = round_up(U1, 32), nthreads = half * (half <= 512 ? 2 : 1);
        if (nthreads <= 256) rnnt_lattice_kernel<256><<<N, nthreads, 0, st>>>(lp);
        else rnnt_lattice_kernel<1024><<<N, nthreads, 0, st>>>(lp);
    }
    return check_launch("rnnt_lattice_kernel");
}

int ha_rnnt_fg_bwd(const float* f, const float* g, int N, int T, int U1, int V,
                   const float* grad_loss, float* gf, float* gg,
                   void* ws, size_t ws_bytes, void* stream) {
    if (U1 <= 0) return fail(HA_ERR_INVALID_ARGUMENT, "U1 must be >= 1");
    const RnntFgWs w = rnnt_fg_ws_layout(N, T, U1);
    int rc = common_checks(f, T, N, V, U1 - 1, ws, ws_bytes, ha_rnnt_fg_workspace_bytes(N, T, U1, V));
    if (rc) return rc;
    if (!g || !grad_loss || !gf || !gg) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;
    RnntFgParams p = rnnt_fg_params(f, g, N, T, U1, V, w, base);
    p.gf = gf; p.gg = gg; p.gout = grad_loss;
    const int nt = fg_umma_nt(V);
    if (nt && aligned16(f) && aligned16(g) && aligned16(gf) && aligned16(gg)) {
        // (the forward call took the same branch: it depends on V and on the alignment of f and g only)
        const FgUmmaWs uw = fg_umma_ws_layout(N, T, U1, V);
        const FgUmmaParams q = fg_umma_params(p, uw, base + w.total);
        fg_w_kernel<<<dim3((uw.Tp + 7) / 8, N), 256, 0, st>>>(q);
        if ((rc = check_launch("fg_w_kernel"))) return rc;
        // DF = (W G) (.) F: rows of W (Tp x Uk, K-major) against G (Um x V) read as it lies (MN-major: the contraction runs over
        // its rows u), K = Uk.  DG = (W^T F) (.) G: W (Tp x Uk) and F (Tp x V) both as they lie (the contraction runs over
        // their rows t), K = Tk.  No transposed copies of F, G or W exist.
        if ((rc = fg_engine_launch<kUmmaDF>({q.Wh, (size_t)uw.Tp, (size_t)uw.Uk, 0}, {q.Gh, (size_t)uw.Um, (size_t)V, 1}, N, uw.Tp / kHM, V, uw.Uk, q, st,
                                            "umma_gemm_kernel<fg DF>"))) return rc;
        if ((rc = fg_engine_launch<kUmmaDG>({q.Wh, (size_t)uw.Tp, (size_t)uw.Uk, 1}, {q.Fh, (size_t)uw.Tp, (size_t)V, 1}, N, uw.Um / kHM, V, uw.Tk, q, st,
                                            "umma_gemm_kernel<fg DG>"))) return rc;
    } else {
        rnnt_fg_gemm_kernel<kDF><<<dim3((T + kGM - 1) / kGM, (V + kGN - 1) / kGN, N), 256, 0, st>>>(p);
        if ((rc = check_launch("rnnt_fg_gemm_kernel<DF>"))) return rc;
        rnnt_fg_gemm_kernel<kDG><<<dim3((U1 + kGM - 1) / kGM, (V + kGN - 1) / kGN, N), 256, 0, st>>>(p);
        if ((rc = check_launch("rnnt_fg_gemm_kernel<DG>"))) return rc;
    }
    rnnt_fg_fix_kernel<<<dim3((T + U1 + 7) / 8, N), 256, 0, st>>>(p);
    return check_launch("rnnt_fg_fix_kernel");
}

// ------------------------------------------------------------------------------- alignment ---
int ha_greedy_decode(const float* x, int64_t sx_n, int64_t sx_t, int N, int T, int V,
                     const void* in_len, int lengths_i64,
                     int64_t* alignment, float* score, int64_t* hyp, int64_t* hyp_len, void* stream) {
    if (!x || !alignment || !score || !hyp || !hyp_len) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (N <= 0 || T <= 0 || V <= 0 || N > 65535) return fail(HA_ERR_INVALID_ARGUMENT, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    GreedyParams p{};
    p.x = x; p.sx_n = sx_n; p.sx_t = sx_t; p.N = N; p.T = T; p.V = V;
    p.in_len = in_len; p.len64 = lengths_i64;
    p.alignment = (long long*)alignment; p.score = score; p.hyp = (long long*)hyp; p.hyp_len = (long long*)hyp_len;
    greedy_argmax_kernel<<<dim3((T + 7) / 8, N), 256, 0, st>>>(p);
    int rc = check_launch("greedy_argmax_kernel");
    if (rc) return rc;
    greedy_collapse_kernel<<<N, 256, 0, st>>>(p);
    return check_launch("greedy_collapse_kernel");
}

size_t ha_ctc_beam_search_workspace_bytes(int N, int T, int V, int beam) {
    if (N <= 0 || T <= 0 || V <= 0 || beam <= 0 || beam > kMaxBeam) return 0;
    return beam_ws_bytes(N, T, V, beam);
}

int ha_ctc_beam_search(const float* lp, int64_t sx_n, int64_t sx_t, int N, int T, int V,
                       const void* in_len, int lengths_i64, int beam, int reference_ext_blank,
                       int64_t* hyp, int64_t* hyp_len, float* score, void* ws, size_t ws_bytes, void* stream) {
    if (!lp || !hyp || !hyp_len || !score || !ws) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    if (N <= 0 || T <= 0 || V <= 0 || N > 65535) return fail(HA_ERR_INVALID_ARGUMENT, "bad sizes");
    if (beam < 1 || beam > kMaxBeam) return fail(HA_ERR_UNSUPPORTED_SHAPE, "beam size must be 1..%d", kMaxBeam);
    if (V > 65534) return fail(HA_ERR_UNSUPPORTED_SHAPE, "V > 65534 is not supported by the beam search");
    if (ws_bytes < beam_ws_bytes(N, T, V, beam)) return fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, beam_ws_bytes(N, T, V, beam));
    BeamParams p{};
    p.lp = lp; p.sx_n = sx_n; p.sx_t = sx_t; p.N = N; p.T = T; p.V = V; p.beam = beam;
    p.in_len = in_len; p.len64 = lengths_i64;
    p.ext_blank = reference_ext_blank ? 0.0f : -HUGE_VALF;
    p.bp = (int*)ws;
    p.cand = (float*)((unsigned char*)ws + round_up_sz((size_t)N * T * beam * 4, 256));
    p.hyp = (long long*)hyp; p.hyp_len = (long long*)hyp_len; p.score = score;
    ctc_beam_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("ctc_beam_kernel");
}

size_t ha_ctc_viterbi_workspace_bytes(int T, int N, int V, int S) {
    (void)V;
    if (T <= 0 || N <= 0 || S < 0) return 0;
    return viterbi_ws_layout(T, N, S).total;
}

int ha_ctc_viterbi(const float* lp, int64_t sx_t, int64_t sx_n, int T, int N, int V,
                   const void* targets, int64_t tgt_stride, int S, int targets_i64,
                   const void* in_len, const void* tgt_len, int lengths_i64,
                   int64_t* alignment, float* score, void* ws, size_t ws_bytes, void* stream) {
    const ViterbiWs w = viterbi_ws_layout(T, N, S);
    int rc = common_checks(lp, T, N, V, S, ws, ws_bytes, w.total);
    if (rc) return rc;
    if (!in_len || !tgt_len || !alignment || !score || (S > 0 && !targets)) return fail(HA_ERR_INVALID_ARGUMENT, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)ws;
    PrepParams pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = w.Sp;
    pp.meta = (int4*)(base + w.meta); pp.order = (int*)(base + w.order);
    pp.tgt = (int*)(base + w.tgt); pp.dupnext = (int*)(base + w.dupnext); pp.star = 0;
    ctc_prep_kernel<<<N, 256, (size_t)round_up(S > 0 ? S : 1, 4) * 4, st>>>(pp);
    if ((rc = check_launch("ctc_prep_kernel"))) return rc;
    ViterbiParams vp{};
    vp.lp = lp; vp.sx_t = sx_t; vp.sx_n = sx_n; vp.T = T; vp.N = N; vp.V = V; vp.S_ = w.S_;
    vp.meta = pp.meta; vp.tgt = pp.tgt; vp.Sp = w.Sp; vp.bp = base + w.bp;
    vp.alignment = (long long*)alignment; vp.score = score;
    const size_t smem = (size_t)w.S_ * 12;
    if ((rc = set_smem(ctc_viterbi_kernel, smem, "ctc_viterbi"))) return rc;
    ctc_viterbi_kernel<<<N, 256, smem, st>>>(vp);
    return check_launch("ctc_viterbi_kernel");
}

}  // extern "C"

// star2.cu — host side of the fused star-CTC path (star2.cuh): eligibility, workspace, launches.
#include <cmath>
#include <cstdlib>

#include "../../include/ha_b200.h"
#include "host.h"
#include "star2.cuh"

namespace hab {

namespace {

struct CfgS { int W, R, J, NS; size_t smem_fwd, smem_bwd; bool ok; };

int env_int(const char* name) {
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}

// Quads per lane J and trellis warps W for the longest target the batch may hold (fewer quads per lane = more warps
// sharing a step's arithmetic), row warps R and the ring depth of each.  The backward kernel wants rows in flight (a
// row occupies its stage from the load until its gradient has been stored): 8 row warps with 3 stages where two CTAs
// still fit an SM, then fewer.
CfgS pick_cfg_s(int V, int S) {
    CfgS c{0, 0, 0, 0, 0, 0, false};
    if (V % 4 != 0 || V < 4) return c;
    const int Sp = round_up(S > 0 ? S : 1, 4);
    const int NLmax = S / 4 + 2, NAmax = 4 * NLmax;
    static const int force_r = env_int("HA_B200_STAR_R"), force_ns = env_int("HA_B200_STAR_NS"),
                     force_j = env_int("HA_B200_STAR_J");                                              // tuning only
    // instantiated (J, W) pairs, most parallel first
    static const struct { int J, W; } kJW[] = {{2, 2}, {2, 4}, {2, 5}, {4, 1}, {4, 2}, {4, 3}, {4, 5}};
    static const struct { int J, W; } kJW1[] = {{1, 4}, {1, 7}};
    auto fits = [&](int J, int W) { return (NAmax / J + 31) / 32 <= W; };
    if (force_j == 1) for (const auto& t : kJW1) if (!c.W && fits(t.J, t.W)) { c.J = t.J; c.W = t.W; }
    for (const auto& t : kJW) {
        if (c.W) break;
        if (force_j && t.J != force_j) continue;
        if (fits(t.J, t.W)) { c.J = t.J; c.W = t.W; }
    }
    if (!c.W) for (const auto& t : kJW) if (!c.W && fits(t.J, t.W)) { c.J = t.J; c.W = t.W; }
    if (!c.W) return c;
    static const struct { int R, NS; size_t cap; } kTry[] = {
        {8, 3, 112 * 1024}, {4, 3, 112 * 1024}, {4, 2, 112 * 1024}, {4, 2, 200 * 1024}, {4, 1, 224 * 1024}};
    for (const auto& t : kTry) {
        if (c.ok) break;
        int R = t.R, NS = t.NS; size_t cap = t.cap;
        if ((force_r == 4 || force_r == 8) && force_ns >= 1 && force_ns <= 4) { R = force_r; NS = force_ns; cap = 224 * 1024; }
        if (c.J == 1 && R != 8) continue;
        const size_t f = star2_smem(c.W, R, NS, V, Sp, NLmax, false).total, b = star2_smem(c.W, R, NS, V, Sp, NLmax, true).total;
        if (b <= cap) { c.R = R; c.NS = NS; c.smem_fwd = f; c.smem_bwd = b; c.ok = true; }
    }
    if (!c.ok && c.J == 1) { c.J = 0; c.W = 0; }
    return c;
}

bool force_legacy_star() {
    static const bool v = [] { const char* e = getenv("HA_B200_STAR_LEGACY"); return e && e[0] == '1'; }();
    return v;
}

template <typename K>
int set_smem_s(K kernel, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return host_fail(HA_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return HA_OK;
}

Star2Params base_params(const Star2Ws& w, unsigned char* base, int T, int N, int V, int S, const CfgS& c) {
    Star2Params p{};
    p.T = T; p.N = N; p.V = V; p.S = S; p.Sp = w.Sp;
    p.meta = (const int4*)(base + w.meta); p.order = (const int*)(base + w.order);
    p.tgt = (const int*)(base + w.tgt); p.nflist = (const int*)(base + w.nflist); p.nfhdr = (const int2*)(base + w.nfhdr); p.NF = w.NF;
    p.stat = (float4*)(base + w.stat); p.tr = (int*)(base + w.tr); p.SPL = w.SPL;
    p.bound = (int*)(base + w.bound); p.BW = w.BW; p.scr = (int*)(base + w.scr); p.zinfo = (int4*)(base + w.zinfo); p.cnt = (int*)(base + w.cnt);
    p.loss_ws = (float*)(base + w.loss); p.hdr = (float*)base;
    p.NS = c.NS; p.NLmax = w.NLmax; p.NA = 4 * w.NLmax; p.EMF = star2_em_floats(w.NLmax);
    return p;
}

// (W, R, J, min CTAs per SM)
#define HAB_SW_CASES(X)                                                                                     \
    X(1, 4, 4, 4) X(2, 4, 4, 3) X(3, 4, 4, 3) X(5, 4, 4, 2) X(1, 8, 4, 2) X(2, 8, 4, 2) X(3, 8, 4, 2) X(5, 8, 4, 1) \
    X(2, 4, 2, 3) X(4, 4, 2, 2) X(5, 4, 2, 2) X(2, 8, 2, 2) X(4, 8, 2, 2) X(5, 8, 2, 1) X(4, 8, 1, 2) X(7, 8, 1, 2)

}  // namespace

bool star2_eligible(int T, int N, int V, int S) {
    if (T <= 0 || N <= 0 || S < 0 || S > 511 || V < 4 || force_legacy_star()) return false;
    return pick_cfg_s(V, S).ok;
}

size_t star2_workspace_bytes(int T, int N, int S) { return star2_ws_layout(T, N, S).total; }

int star2_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
              const void* targets, int64_t tgt_stride, int S, int targets_i64,
              const void* in_len, const void* tgt_len, int lengths_i64,
              float star_penalty, int from_logits, float* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Star2Ws w = star2_ws_layout(T, N, S);
    if (ws_bytes < w.total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, w.total);
    if (!aligned16(x) || (sx_t % 4) || (sx_n % 4))
        return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_star_fwd: x must be 16-byte aligned with strides that are multiples of 4 elements");
    const CfgS c = pick_cfg_s(V, S);
    unsigned char* base = (unsigned char*)ws;
    int rc;
    Prep2Params pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = w.Sp;
    pp.meta = (int4*)(base + w.meta); pp.order = (int*)(base + w.order);
    pp.tgt = (int*)(base + w.tgt); pp.nflist = (int*)(base + w.nflist); pp.nfhdr = (int2*)(base + w.nfhdr); pp.NF = w.NF;
    pp.cnt = (int*)(base + w.cnt); pp.zinfo = (int4*)(base + w.zinfo); pp.star = 1;
    ctc2_prep_kernel<<<N, 256, (size_t)w.Sp * 8, st>>>(pp);
    if ((rc = host_check_launch("ctc2_prep_kernel"))) return rc;

    Star2Params p = base_params(w, base, T, N, V, S, c);
    p.x = x; p.sx_t = sx_t; p.sx_n = sx_n; p.loss = loss; p.from_logits = from_logits;
    // exp(star_penalty), kept inside the range where a product with a normal number stays exact enough (the penalty
    // multiplies the star states every frame; beyond e^-69 a star is never worth taking)
    p.pen = (float)std::exp(std::fmin(std::fmax((double)star_penalty, -69.0), 40.0));
    const dim3 grid(2 * N);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    bool hit = false;
#define HAB_CASE(WW, RR, JJ, MB)                                                                            \
    if (!hit && c.W == WW && c.R == RR && c.J == JJ) {                                                         \
        hit = true;                                                                                   \
        static bool attr[64] = {};                                                                    \
        if (!attr[dev]) { if ((rc = set_smem_s(star2_fwd_kernel<WW, RR, JJ, MB>, kMaxSmemOptin, "star2_fwd"))) return rc; attr[dev] = true; } \
        star2_fwd_kernel<WW, RR, JJ, MB><<<grid, 32 * (WW + RR), c.smem_fwd, st>>>(p);                        \
    }
    HAB_SW_CASES(HAB_CASE)
#undef HAB_CASE
    if (!hit) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: star2 W=%d R=%d J=%d", c.W, c.R, c.J);
    return host_check_launch("star2_fwd_kernel");
}

int star2_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
              const float* grad_loss, int from_logits, float* gx, int64_t sg_t, int64_t sg_n,
              void* ws, size_t ws_bytes, cudaStream_t st) {
    const Star2Ws w = star2_ws_layout(T, N, S);
    if (ws_bytes < w.total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, w.total);
    if (!x) return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_star_bwd: x is required (the emissions are re-gathered from it)");
    if (!aligned16(x) || (sx_t % 4) || (sx_n % 4) || !aligned16(gx) || (sg_t % 4) || (sg_n % 4))
        return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_star_bwd: x and gx must be 16-byte aligned with strides that are multiples of 4 elements");
    const CfgS c = pick_cfg_s(V, S);
    unsigned char* base = (unsigned char*)ws;
    int rc;
    Star2Params p = base_params(w, base, T, N, V, S, c);
    p.x = x; p.sx_t = sx_t; p.sx_n = sx_n; p.gx = gx; p.sg_t = sg_t; p.sg_n = sg_n;
    p.gout = grad_loss; p.from_logits = from_logits;
    const dim3 grid(2 * N);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    bool hit = false;
#define HAB_CASE(WW, RR, JJ, MB)                                                                            \
    if (!hit && c.W == WW && c.R == RR && c.J == JJ) {                                                         \
        hit = true;                                                                                   \
        static bool attr[64] = {};                                                                    \
        if (!attr[dev]) { if ((rc = set_smem_s(star2_bwd_kernel<WW, RR, JJ, MB>, kMaxSmemOptin, "star2_bwd"))) return rc; attr[dev] = true; } \
        star2_bwd_kernel<WW, RR, JJ, MB><<<grid, 32 * (WW + RR), c.smem_bwd, st>>>(p);                        \
    }
    HAB_SW_CASES(HAB_CASE)
#undef HAB_CASE
    if (!hit) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: star2 W=%d R=%d J=%d", c.W, c.R, c.J);
    return host_check_launch("star2_bwd_kernel");
}

}  // namespace hab

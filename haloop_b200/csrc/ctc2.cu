// ctc2.cu — host side of the fused CTC path (ctc2.cuh): eligibility, workspace, launches.
#include <cstdlib>

#include "../../include/ha_b200.h"
#include "ctc2.cuh"
#include "host.h"

namespace hab {

namespace {

struct Cfg2 { int W, NS; size_t smem_fwd, smem_bwd; bool ok; };

// trellis warps for the longest target the batch may hold, ring depth of the row warps
Cfg2 pick_cfg2(int V, int S) {
    Cfg2 c{0, 0, 0, 0, false};
    const int Sp = round_up(S > 0 ? S : 1, 4);
    const int NLmax = Sp / 4 + 2;
    const int Wn = (NLmax + 31) / 32;
    static const int kW[] = {1, 2, 3, 5, 9};
    for (int w : kW) if (!c.W && w >= Wn) c.W = w;
    if (!c.W || V % 4 != 0 || V < 4) return c;
    const int SPL = 8 * NLmax;
    for (int ns = 2; ns >= 1 && !c.ok; --ns) {
        const size_t f = ctc2_smem(c.W, ns, V, Sp, SPL, false).total, b = ctc2_smem(c.W, ns, V, Sp, SPL, true).total;
        const size_t cap = (ns == 2) ? 56 * 1024 : 110 * 1024;       // 4 CTAs per SM with two stages, 2 with one
        if (b <= cap) { c.NS = ns; c.smem_fwd = f; c.smem_bwd = b; c.ok = true; }
    }
    return c;
}

bool force_legacy() {
    static const bool v = [] { const char* e = getenv("HA_B200_CTC_LEGACY"); return e && e[0] == '1'; }();
    return v;
}

template <typename K>
int set_smem2(K kernel, size_t bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return host_fail(HA_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return HA_OK;
}

Ctc2Params base_params(const Ctc2Ws& w, unsigned char* base, int T, int N, int V, const Cfg2& c) {
    Ctc2Params p{};
    p.T = T; p.N = N; p.V = V; p.Sp = w.Sp;
    p.meta = (const int4*)(base + w.meta); p.order = (const int*)(base + w.order);
    p.tgt = (const int*)(base + w.tgt); p.nflist = (const int*)(base + w.nflist); p.nfhdr = (const int2*)(base + w.nfhdr); p.NF = w.NF;
    p.lse2 = (float*)(base + w.lse2); p.tr = (int*)(base + w.tr); p.SPL = w.SPL;
    p.bound = (int*)(base + w.bound); p.BW = w.BW; p.zinfo = (int4*)(base + w.zinfo); p.cnt = (int*)(base + w.cnt);
    p.loss_ws = (float*)(base + w.loss);
    p.NS = c.NS; p.NLmax = w.NLmax; p.EMF = ctc2_em_floats(w.Sp);
    return p;
}

#define HAB_W_CASES(X) X(1, 6) X(2, 5) X(3, 4) X(5, 3) X(9, 2)

}  // namespace

bool ctc2_eligible(int T, int N, int V, int S) {
    if (T <= 0 || N <= 0 || S < 0 || S > 1023 || force_legacy()) return false;
    return pick_cfg2(V, S).ok;
}

size_t ctc2_workspace_bytes(int T, int N, int S) { return ctc2_ws_layout(T, N, S).total; }

int ctc2_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
             const void* targets, int64_t tgt_stride, int S, int targets_i64,
             const void* in_len, const void* tgt_len, int lengths_i64,
             int from_logits, float* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
    const Ctc2Ws w = ctc2_ws_layout(T, N, S);
    if (ws_bytes < w.total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, w.total);
    if (!aligned16(x) || (sx_t % 4) || (sx_n % 4))
        return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_ctc_fwd: x must be 16-byte aligned with strides that are multiples of 4 elements");
    const Cfg2 c = pick_cfg2(V, S);
    unsigned char* base = (unsigned char*)ws;
    int rc;
    Prep2Params pp{};
    pp.targets = targets; pp.tgt_stride = tgt_stride; pp.tgt64 = targets_i64;
    pp.in_len = in_len; pp.tgt_len = tgt_len; pp.len64 = lengths_i64;
    pp.T = T; pp.N = N; pp.V = V; pp.S = S; pp.Sp = w.Sp;
    pp.meta = (int4*)(base + w.meta); pp.order = (int*)(base + w.order);
    pp.tgt = (int*)(base + w.tgt); pp.nflist = (int*)(base + w.nflist); pp.nfhdr = (int2*)(base + w.nfhdr); pp.NF = w.NF;
    pp.cnt = (int*)(base + w.cnt); pp.zinfo = (int4*)(base + w.zinfo);
    ctc2_prep_kernel<<<N, 256, (size_t)w.Sp * 8, st>>>(pp);
    if ((rc = host_check_launch("ctc2_prep_kernel"))) return rc;

    Ctc2Params p = base_params(w, base, T, N, V, c);
    p.x = x; p.sx_t = sx_t; p.sx_n = sx_n; p.loss = loss; p.from_logits = from_logits;
    const dim3 grid(2 * N);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    bool hit = false;
#define HAB_CASE(WW, MB)                                                                              \
    if (!hit && c.W == WW) {                                                                          \
        hit = true;                                                                                   \
        static bool attr[64] = {};                                                                    \
        if (!attr[dev]) { if ((rc = set_smem2(ctc2_fwd_kernel<WW, MB>, kMaxSmemOptin, "ctc2_fwd"))) return rc; attr[dev] = true; } \
        ctc2_fwd_kernel<WW, MB><<<grid, 32 * (WW + kR2), c.smem_fwd, st>>>(p);                         \
    }
    HAB_W_CASES(HAB_CASE)
#undef HAB_CASE
    if (!hit) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: ctc2 W=%d", c.W);
    return host_check_launch("ctc2_fwd_kernel");
}

int ctc2_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
             const float* grad_loss, int from_logits, float* gx, int64_t sg_t, int64_t sg_n,
             void* ws, size_t ws_bytes, cudaStream_t st) {
    const Ctc2Ws w = ctc2_ws_layout(T, N, S);
    if (ws_bytes < w.total) return host_fail(HA_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < %zu", ws_bytes, w.total);
    if (!x) return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_ctc_bwd: x is required (the emissions are re-gathered from it)");
    if (!aligned16(x) || (sx_t % 4) || (sx_n % 4) || !aligned16(gx) || (sg_t % 4) || (sg_n % 4))
        return host_fail(HA_ERR_INVALID_ARGUMENT, "ha_ctc_bwd: x and gx must be 16-byte aligned with strides that are multiples of 4 elements");
    const Cfg2 c = pick_cfg2(V, S);
    unsigned char* base = (unsigned char*)ws;
    int rc;
    Ctc2Params p = base_params(w, base, T, N, V, c);
    p.x = x; p.sx_t = sx_t; p.sx_n = sx_n; p.gx = gx; p.sg_t = sg_t; p.sg_n = sg_n;
    p.gout = grad_loss; p.from_logits = from_logits;
    const dim3 grid(2 * N);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    bool hit = false;
#define HAB_CASE(WW, MB)                                                                              \
    if (!hit && c.W == WW) {                                                                          \
        hit = true;                                                                                   \
        static bool attr[64] = {};                                                                    \
        if (!attr[dev]) { if ((rc = set_smem2(ctc2_bwd_kernel<WW, MB>, kMaxSmemOptin, "ctc2_bwd"))) return rc; attr[dev] = true; } \
        ctc2_bwd_kernel<WW, MB><<<grid, 32 * (WW + kR2), c.smem_bwd, st>>>(p);                         \
    }
    HAB_W_CASES(HAB_CASE)
#undef HAB_CASE
    if (!hit) return host_fail(HA_ERR_UNSUPPORTED_SHAPE, "internal: ctc2 W=%d", c.W);
    return host_check_launch("ctc2_bwd_kernel");
}

}  // namespace hab

// common.cuh — sm_100a device helpers shared by the CTC / star-CTC / RNN-T kernels.
//
// Log-domain quantities (row log-sum-exps, emission exponents) are kept in LOG2 units; the recursions
// themselves run in the linear domain on extended-range numbers (XF below).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace hab {

constexpr float kVoid = -1.0e30f;       // "void" (ha/ctc.py:135 uses finfo.min); finite so void-void = 0, never NaN
constexpr float kVoidTest = -1.0e29f;   // anything below this is void
constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453094;
constexpr double kLog2e_d = 1.4426950408889634074;

constexpr int kLabelMask = 0x3fffffff;  // tgt words: label | kNotFirst
constexpr int kNotFirst = 0x40000000;   // an earlier position of the same utterance holds the same label

__host__ __device__ constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }
__host__ __device__ constexpr size_t round_up_sz(size_t v, size_t m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: (x + kMagic) - kMagic == rint(x) for |x| < 2^22 (F2I-free)
__device__ __forceinline__ float round_int(float x) { return __fsub_rn(__fadd_rn(x, kMagic), kMagic); }
// ---- emissions in float-float arithmetic ----------------------------------------------------------
// e = x * log2(e) - l2 - ct for an fp32 logit x, split into an integer part K (clamped at -127) and a
// fraction f with |error| ~1e-8: two-product and two-sum error-free transformations on the FP32 pipe
// (float64 conversions run at 1/16 rate and made the row kernels compute bound).
constexpr float kLog2eHi = 1.4426950216293335f;             // fl(log2 e)
constexpr float kLog2eLo = 1.9259629911266175e-8f;          // log2 e - fl(log2 e)
__device__ __forceinline__ void emission_split(float x, float l2, float ct, float& K, float& f) {
    const float ph = x * kLog2eHi;
    const float pl = fmaf(x, kLog2eLo, fmaf(x, kLog2eHi, -ph));        // x*log2e = ph + pl
    const float s = ph - l2;                                             // two-sum of ph + (-l2)
    const float bb = s - ph;
    const float err = (ph - (s - bb)) + (-l2 - bb);
    const float Kt = fmaxf(round_int(s - ct) + ct, ct - 127.0f);         // integer part of the unshifted value
    K = Kt - ct;
    f = (s - Kt) + (err + pl);
}

// ---- extended-range linear numbers ----------------------------------------------------------------
// value = m * 2^e with m an fp32 in [1, 2) (or a small multiple of it between a sum and the next
// normalisation) and e an int32.  The CTC trellis runs in this representation: the sum-product
// recursion costs integer exponent alignment + one FADD + one FMUL per transition instead of a
// log-add-exp (2 MUFU + ~15 FP32 on the split log-domain numbers this replaced), keeps 24 significant bits whatever the magnitude, and
// has no range limit (the exponent is a full int).  "void" (probability 0) is m = 1, e = kVoidE: it
// aligns to +0 against anything real, and adding voids keeps the exponent far below kVoidETest.
struct XF { float m; int e; };
constexpr int kVoidE = -(1 << 28);
constexpr int kVoidETest = -(1 << 27);
__device__ __forceinline__ XF xf_make(float m, int e) { XF r; r.m = m; r.e = e; return r; }
__device__ __forceinline__ float xf_scale(float m, int d) {      // m * 2^d by exponent-field arithmetic
    return __int_as_float(__float_as_int(m) + (int)((unsigned)d << 23));
}
// a + b, not normalised: mantissa < 2 max(ma, mb) at exponent max(ea, eb).  A term more than 2^127 below
// the other one drops to (at most) a denormal.  Mantissas must be >= 1.
__device__ __forceinline__ XF xf_add(XF a, XF b) {
    const int ex = max(a.e, b.e);
    const float xa = xf_scale(a.m, max(a.e - ex, -127));
    const float xb = xf_scale(b.m, max(b.e - ex, -127));
    return xf_make(xa + xb, ex);
}
// (s * p), mantissa back in [1, 2).  s.m * p must be a positive normal float.
__device__ __forceinline__ XF xf_mul_norm(XF s, float p) {
    const int rb = __float_as_int(s.m * p);
    return xf_make(__int_as_float((rb & 0x007fffff) | 0x3f800000), s.e + (rb >> 23) - 127);
}
// 2^f for |f| <= 0.5: degree-7 polynomial on the FP32 pipe, relative error ~1e-7
__device__ __forceinline__ float exp2_poly(float f) {
    float r = 1.5252733804059841e-05f;
    r = fmaf(r, f, 1.5403530393381608e-04f);
    r = fmaf(r, f, 1.3333558146428443e-03f);
    r = fmaf(r, f, 9.618129107628477e-03f);
    r = fmaf(r, f, 5.550410866482158e-02f);
    r = fmaf(r, f, 2.402265069591007e-01f);
    r = fmaf(r, f, 6.931471805599453e-01f);
    return fmaf(r, f, 1.0f);
}

// 2^(K + f) for an integer-valued K in [-127, 1] and |f| <= 0.5 (exp2_poly: relative error ~1e-7, no
// MUFU), exponent added into the bit pattern
__device__ __forceinline__ float emission_linear(float K, float f) {
    const int k = __float_as_int(K + kMagic) - 0x4B400000;          // K is integer valued: exact, no F2I
    return (k < -125) ? 1.1754943508222875e-38f : xf_scale(exp2_poly(f), k);
}

// stored trellis word: [12 bits: exponent distance below the slot base, saturating][20 mantissa bits]
constexpr unsigned kPackVoid = 0xfff00000u;
__device__ __forceinline__ int xf_pack(float m, int below) {     // m in [1,2), below >= 0
    return (int)__funnelshift_r((unsigned)__float_as_int(m) << 9, (unsigned)min(below, 4095), 12);
}
__device__ __forceinline__ float xf_unpack_m(int w) {            // mantissa, truncation re-centred
    return __int_as_float((int)((((unsigned)w << 3) & 0x007ffff8u) | 0x3f800004u));
}
__device__ __forceinline__ int xf_unpack_below(int w) { return (int)((unsigned)w >> 20); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// ---- mbarrier + bulk async copy (TMA engine; SASS: UBLKCP / SYNCS) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)     // suspend-time hint: the wait sleeps in hardware
        : "memory");
    return ok != 0;
}
// non-blocking phase test (the result can be consumed many instructions later: the latency of the test overlaps them)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 26); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// Wait of a warp that is not on the critical path (a row warp waiting for the trellis warps): sleep between
// polls so that it leaves the issue slots to the warps that have work.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        __nanosleep(128);
        if (mbar_try_wait(bar, parity)) return;
    }
    __trap();
}
// Producer-side wait: back off between polls so a spinning producer lane does not take issue slots
// from the compute warps of its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(1000);
    }
    __trap();
}
// global -> shared, completion counted on an mbarrier. bytes % 16 == 0, both addresses 16B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                     __cvta_generic_to_global(dst_gmem)),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// order this thread's generic-proxy writes before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// non-blocking arrival on a named barrier (the waiting side uses named_bar_sync with the same count)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// The phase-switch barrier of the trellis kernels: producer and compute warps come from different code
// paths, so the barrier instruction lives in one non-inlined function and every thread of the CTA
// executes the same instruction (bar.sync on a named barrier with an explicit thread count).
static __device__ __noinline__ void cta_phase_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// index loads for int32 / int64 targets and lengths (the reference passes LongTensors from
// Collator, ha/loop.py:29-41, and int32 in its own tests, ha/transducer.py:217-218)
__device__ __forceinline__ long long load_idx(const void* p, long long i, int is64) {
    return is64 ? ((const long long*)p)[i] : (long long)((const int*)p)[i];
}

}  // namespace hab

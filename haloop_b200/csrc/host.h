// host.h — host-side helpers shared by the translation units of libha_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace hab {

// records the thread-local message ha_b200_last_error() returns; returns `code`
int host_fail(int code, const char* fmt, ...);
int host_check_launch(const char* what);

constexpr size_t kMaxSmemOptin = 227 * 1024;       // opt-in dynamic shared memory per CTA on sm_100

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- fused CTC path (ctc2.cu) ----
bool ctc2_eligible(int T, int N, int V, int S);
size_t ctc2_workspace_bytes(int T, int N, int S);
int ctc2_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
             const void* targets, int64_t tgt_stride, int S, int targets_i64,
             const void* in_len, const void* tgt_len, int lengths_i64,
             int from_logits, float* loss, void* ws, size_t ws_bytes, cudaStream_t st);
int ctc2_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
             const float* grad_loss, int from_logits, float* gx, int64_t sg_t, int64_t sg_n,
             void* ws, size_t ws_bytes, cudaStream_t st);

// ---- fused star-CTC path (star2.cu) ----
bool star2_eligible(int T, int N, int V, int S);
size_t star2_workspace_bytes(int T, int N, int S);
int star2_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
              const void* targets, int64_t tgt_stride, int S, int targets_i64,
              const void* in_len, const void* tgt_len, int lengths_i64,
              float star_penalty, int from_logits, float* loss, void* ws, size_t ws_bytes, cudaStream_t st);
int star2_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
              const float* grad_loss, int from_logits, float* gx, int64_t sg_t, int64_t sg_n,
              void* ws, size_t ws_bytes, cudaStream_t st);

// ---- classifier head + CTC (head.cu) reuses the CTC prep and trellis kernels instantiated in api.cu ----
void ctc_head_dims(int S, int* Sp, int* E, int* JWp, int* SPX);
int ctc_prep_for_head(const void* targets, int64_t tgt_stride, int S, int targets_i64,
                      const void* in_len, const void* tgt_len, int lengths_i64, int T, int N, int V, int Sp,
                      void* meta, int* order, int* tgt, int* dupnext, cudaStream_t st);
int ctc_trellis_for_head(int T, int N, int S, int Sp, int E, int SPX, int JWp, const void* meta, const int* order,
                         const int* tgt, float* em, float* tr, float* loss, float* loss_ws, cudaStream_t st);

// ---- tensor maps for the GEMM engine (head.cu) ----
// (rows x cols) fp32 matrix, cols contiguous, leading dimension ld floats.  mn = 0: a K-major operand (cols = the
// contraction index), box 128 rows x 32 cols, SWIZZLE_128B; mn = 1: an MN-major operand (rows = the contraction index), box
// 32 x 32, SWIZZLE_128B_ATOM_32B.  `map` points at a CUtensorMap.
int host_make_map(void* map, const float* ptr, size_t rows, size_t cols, size_t ld, int mn = 0);
int host_sm_count();

}  // namespace hab

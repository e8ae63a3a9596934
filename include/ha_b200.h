/*
 * ha_b200.h — C ABI of libha_b200.so: B200 (sm_100a) alignment losses for haloop.
 *
 * The reference (proger/haloop) has no FFI: its hot path is four Python functions imported by
 * name into ha/recognizer.py:6-8.  Each entry point below names the reference function whose
 * arithmetic it replaces; haloop_b200/ops.py binds them with ctypes and registers them as
 * PyTorch custom ops with autograd (INTEGRATION.md shows the binding and the call-site patch).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; no torch types
 *   - the library never allocates, never synchronises and keeps no global state: scratch and
 *     saved-for-backward state live in a caller-owned workspace sized by *_workspace_bytes()
 *   - `stream` is a cudaStream_t passed as void*
 *   - return value: 0 = ok, otherwise an HA_ERR_* code; ha_b200_last_error() (thread-local)
 *     describes it
 *   - logits/log-probs are float32 with unit stride along the class axis; strides are in elements
 *   - targets are 0-padded, blank = 0 (ha/ctc.py:123, ha/star.py:85, ha/transducer.py:190)
 *   - from_logits = 1: input is raw logits, the op applies log-softmax itself and the gradient is
 *     d loss / d logits = (softmax - occupancy) * grad_loss.  from_logits = 0: input is log-probs
 *     used as they are (the reference signature) and the gradient is what autograd yields at the
 *     reference's `emissions` / `joint` argument, -occupancy * grad_loss.
 *   - an infeasible alignment returns loss = +inf and a zero gradient (the reference returns
 *     ~3.4e38, ha/ctc.py:135); an utterance with out-of-range lengths or labels returns NaN.
 */
#ifndef HA_B200_H
#define HA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HA_OK 0
#define HA_ERR_INVALID_ARGUMENT 1
#define HA_ERR_WORKSPACE_TOO_SMALL 2
#define HA_ERR_UNSUPPORTED_SHAPE 3
#define HA_ERR_CUDA 4

int ha_b200_version(void);
const char* ha_b200_last_error(void);

/* ---- CTC: ha/ctc.py:110-174 ctc_forward_score3 (+ its autograd backward) ------------------- */
size_t ha_ctc_workspace_bytes(int T, int N, int V, int S);
/* x (T,N,V) viewed through (sx_t, sx_n, 1); targets (N,S) int32/int64; loss (N) */
int ha_ctc_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
               const void* targets, int64_t tgt_stride, int S, int targets_i64,
               const void* in_len, const void* tgt_len, int lengths_i64,
               int from_logits, float* loss, void* ws, size_t ws_bytes, void* stream);
/* needs the workspace ha_ctc_fwd filled for the same x; grad_loss (N); gx strided like x */
int ha_ctc_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
               const float* grad_loss, int from_logits,
               float* gx, int64_t sg_t, int64_t sg_n,
               void* ws, size_t ws_bytes, void* stream);

/* ---- star-CTC: ha/star.py:65-163 star_ctc_forward_score (+ intersperse_stars :8-49) --------- */
size_t ha_star_workspace_bytes(int T, int N, int V, int S);
int ha_star_fwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V,
                const void* targets, int64_t tgt_stride, int S, int targets_i64,
                const void* in_len, const void* tgt_len, int lengths_i64,
                float star_penalty, int from_logits, float* loss,
                void* ws, size_t ws_bytes, void* stream);
int ha_star_bwd(const float* x, int64_t sx_t, int64_t sx_n, int T, int N, int V, int S,
                const float* grad_loss, int from_logits,
                float* gx, int64_t sg_t, int64_t sg_n,
                void* ws, size_t ws_bytes, void* stream);

/* ---- RNN-T: ha/transducer.py:175-205 transducer_forward_score (+ ha/scan.py:88-126) --------- */
size_t ha_rnnt_workspace_bytes(int N, int T, int U1, int V);
/* joint: a (N,T,U1,V) view with element strides sj_n, sj_t, sj_u and unit class stride, U1 = U+1 (a contiguous joint:
 * T*U1*V, U1*V, V); targets (N,U); gjoint: the gradient, a view of the same shape with strides sg_* */
int ha_rnnt_fwd(const float* joint, int64_t sj_n, int64_t sj_t, int64_t sj_u, int N, int T, int U1, int V,
                const void* targets, int64_t tgt_stride, int targets_i64,
                const void* in_len, const void* tgt_len, int lengths_i64,
                int from_logits, float* loss, void* ws, size_t ws_bytes, void* stream);
int ha_rnnt_bwd(const float* joint, int64_t sj_n, int64_t sj_t, int64_t sj_u, int N, int T, int U1, int V,
                const float* grad_loss, int from_logits, float* gjoint, int64_t sg_n, int64_t sg_t, int64_t sg_u,
                void* ws, size_t ws_bytes, void* stream);

/* ---- joint-free RNN-T: the reference's joint is a broadcast sum, ha/recognizer.py:104-114 ------
 * joint[n,t,u,:] = f[n,t,:] + g[n,u,:] with f = classifier(features) (N,T,V) and g = lm outputs
 * (N,U+1,V), both contiguous raw logits (the log-softmax over the joint is fused).  Same loss as
 * ha_rnnt_fwd on the materialised joint, and the gradients w.r.t. f and g (the reductions of the joint
 * gradient over u and over t), without ever forming the (N,T,U+1,V) tensor or its gradient. */
size_t ha_rnnt_fg_workspace_bytes(int N, int T, int U1, int V);
int ha_rnnt_fg_fwd(const float* f, const float* g, int N, int T, int U1, int V,
                   const void* targets, int64_t tgt_stride, int targets_i64,
                   const void* in_len, const void* tgt_len, int lengths_i64,
                   float* loss, void* ws, size_t ws_bytes, void* stream);
int ha_rnnt_fg_bwd(const float* f, const float* g, int N, int T, int U1, int V,
                   const float* grad_loss, float* gf, float* gg,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- greedy alignment: ha/recognizer.py:48-59 TemporalClassifier.decode -------------------- */
/* x (N,T,V) through (sx_n, sx_t, 1).  alignment (N,T) int64 = per-frame argmax (first index wins,
 * as torch.max), score (N,T) = the max, hyp (N,T) int64 = collapsed repeats with blanks dropped,
 * padded with -1, hyp_len (N) int64.  in_len may be NULL (the reference ignores lengths, :51). */
int ha_greedy_decode(const float* x, int64_t sx_n, int64_t sx_t, int N, int T, int V,
                     const void* in_len, int lengths_i64,
                     int64_t* alignment, float* score, int64_t* hyp, int64_t* hyp_len, void* stream);

/* ---- CTC prefix beam search: ha/beam.py:71-137 ctc_beam_search_decode_logits (the decode the reference leaves
 * commented out at ha/recognizer.py:58 with "FIXME: speed it up"), hypothesis for hypothesis, including its in-place
 * update order, its un-merged duplicate prefixes and the 0.0 blank score of extension candidates.
 * lp (N,T,V) log-probs viewed through (sx_n, sx_t, 1); in_len may be NULL (every frame); beam <= 16.
 * reference_ext_blank = 1: extension candidates enter with blank score 0.0 (log 1) exactly as ha/beam.py:124 does;
 * 0: with -inf (log 0), the algorithm of the reference's probability-domain twin (ha/beam.py:4-68) and of [Graves14].
 * hyp (N,beam,T) int64 padded with -1, best first; hyp_len (N,beam); score (N,beam) = the reference's seq_logits. */
size_t ha_ctc_beam_search_workspace_bytes(int N, int T, int V, int beam);
int ha_ctc_beam_search(const float* lp, int64_t sx_n, int64_t sx_t, int N, int T, int V,
                       const void* in_len, int lengths_i64, int beam, int reference_ext_blank,
                       int64_t* hyp, int64_t* hyp_len, float* score, void* ws, size_t ws_bytes, void* stream);

/* ---- fused classifier head + CTC: replaces, at ha/recognizer.py:43-46 + 61-82,
 *   log_probs = classifier(features).log_softmax(-1);  loss = ctc(log_probs.permute(1,0,2), ...)
 * with one op on the features: h (N,T,D) fp32 contiguous (dropout already applied by the caller), W (V,D) and bias (V)
 * (nn.Linear layout; bias may be NULL).  The (N,T,V) logits, log-probs and their gradient are never written: the
 * GEMM h W^T runs on the tensor cores (tcgen05 kind::tf32) with the log-softmax statistics and the gather of the blank
 * and label logits as its epilogue; the backward recomputes the logits tile by tile and emits dh, dW, db from
 * g (softmax - occupancy) through an L2-sized ring of rows.  precision: 3 = error-compensated tf32 x 3 (fp32-grade,
 * what nn.Linear computes under torch.autocast(dtype=float32)), 1 = plain tf32 (torch's allow_tf32 = True).
 * D must be a multiple of 4 (feature rows are copied by TMA).  `saved` is written by the forward and read by the backward; the scratch buffers
 * are free again when the call's kernels have run.  loss (N), per utterance, as ha_ctc_fwd. */
int ha_head_ctc_workspace_bytes(int N, int T, int D, int V, int S, size_t* saved, size_t* fwd_scratch, size_t* bwd_scratch);
int ha_head_ctc_fwd(const float* h, const float* W, const float* bias, int N, int T, int D, int V,
                    const void* targets, int64_t tgt_stride, int S, int targets_i64,
                    const void* in_len, const void* tgt_len, int lengths_i64, int precision,
                    float* loss, void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, void* stream);
/* dh (N,T,D), dW (V,D), db (V, may be NULL) = gradients of sum_n grad_loss[n] * loss[n] */
int ha_head_ctc_bwd(const float* h, const float* W, const float* bias, int N, int T, int D, int V, int S,
                    const float* grad_loss, int precision, float* dh, float* dW, float* db,
                    void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, void* stream);

/* ---- CTC Viterbi forced alignment (max-semiring of ha/ctc.py:144-167; not in the reference) - */
size_t ha_ctc_viterbi_workspace_bytes(int T, int N, int V, int S);
/* lp (T,N,V) log-probs through (sx_t, sx_n, 1); alignment (N,T) int64 class per frame (-1 beyond
 * the input length); score (N) best-path log-prob */
int ha_ctc_viterbi(const float* lp, int64_t sx_t, int64_t sx_n, int T, int N, int V,
                   const void* targets, int64_t tgt_stride, int S, int targets_i64,
                   const void* in_len, const void* tgt_len, int lengths_i64,
                   int64_t* alignment, float* score, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HA_B200_H */

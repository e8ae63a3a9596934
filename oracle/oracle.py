"""ctypes/numpy front end of oracle/libha_oracle.so. TEST INFRASTRUCTURE ONLY.

Restates, in float64 on the CPU, the reference functions
  ha/ctc.py:110-174 (ctc_forward_score3), ha/ctc.py:177-178 (ctc_reduce_mean),
  ha/star.py:65-163 (star_ctc_forward_score), ha/transducer.py:175-205
  (transducer_forward_score) and ha/recognizer.py:48-59 (greedy decode),
with closed-form gradients.  Pinned against the reference by tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libha_oracle.so")
_lib = None

_f64p = ctypes.POINTER(ctypes.c_double)
_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "ha_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libha_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def ctc(x, targets, in_len, tgt_len, from_logits=True, grad_out=None, want_grad=True):
    """x (T,N,V) -> (loss (N,), grad (T,N,V) or None).  ha/ctc.py:110-174."""
    x = _f64(x); targets = _i64(targets); in_len = _i64(in_len); tgt_len = _i64(tgt_len)
    T, N, V = x.shape
    S = targets.shape[1] if targets.ndim == 2 else 0
    if S == 0:
        targets = np.zeros((N, 1), np.int64); S = 1
    loss = np.empty(N, np.float64)
    grad = np.empty_like(x) if want_grad else None
    go = _f64(grad_out) if grad_out is not None else None
    rc = lib().ha_oracle_ctc(_p(x, _f64p), T, N, V, _p(targets, _i64p), S, _p(in_len, _i64p),
                             _p(tgt_len, _i64p), int(from_logits), _p(go, _f64p),
                             _p(loss, _f64p), _p(grad, _f64p))
    if rc != 0:
        raise ValueError("ha_oracle_ctc: bad lengths")
    return loss, grad


def ctc_reduce_mean(losses, tgt_len):
    """ha/ctc.py:177-178."""
    return float(np.mean(np.asarray(losses, np.float64) / np.asarray(tgt_len, np.float64)))


def head_ctc(h, W, b, targets, in_len, tgt_len, grad_out=None):
    """Classifier head + CTC in float64: ha/recognizer.py:43-46 (`classifier(features).log_softmax(-1)`, dropout
    already applied) followed by ha/ctc.py:110-174, with the gradients of sum_n grad_out[n] loss[n] w.r.t. the
    features h (N,T,D), the weight W (V,D) and the bias b (V) (chain rule through the linear layer).
    -> (loss (N,), dh, dW, db)."""
    h = _f64(h); W = _f64(W)
    b = _f64(b) if b is not None else np.zeros(W.shape[0])
    logits = h @ W.T + b                                     # (N,T,V)
    loss, g = ctc(np.ascontiguousarray(logits.transpose(1, 0, 2)), targets, in_len, tgt_len, True, grad_out)
    g = np.nan_to_num(g.transpose(1, 0, 2), nan=0.0, posinf=0.0, neginf=0.0)      # (N,T,V)
    return loss, g @ W, np.einsum("ntv,ntd->vd", g, h), g.sum((0, 1))


def star(x, targets, in_len, tgt_len, star_penalty=-0.5, from_logits=True, grad_out=None,
         want_grad=True):
    """x (T,N,V) -> (loss (N,), grad).  ha/star.py:65-163."""
    x = _f64(x); targets = _i64(targets); in_len = _i64(in_len); tgt_len = _i64(tgt_len)
    T, N, V = x.shape
    S = targets.shape[1]
    loss = np.empty(N, np.float64)
    grad = np.empty_like(x) if want_grad else None
    go = _f64(grad_out) if grad_out is not None else None
    rc = lib().ha_oracle_star(_p(x, _f64p), T, N, V, _p(targets, _i64p), S, _p(in_len, _i64p),
                              _p(tgt_len, _i64p), ctypes.c_double(star_penalty), int(from_logits),
                              _p(go, _f64p), _p(loss, _f64p), _p(grad, _f64p))
    if rc != 0:
        raise ValueError("ha_oracle_star: bad lengths")
    return loss, grad


def rnnt(joint, targets, in_len, tgt_len, from_logits=True, grad_out=None, want_grad=True):
    """joint (N,T,U+1,V) -> (loss (N,), grad).  ha/transducer.py:175-205."""
    joint = _f64(joint); targets = _i64(targets); in_len = _i64(in_len); tgt_len = _i64(tgt_len)
    N, T, U1, V = joint.shape
    if targets.shape[1] != U1 - 1:
        raise ValueError("targets must be (N, U)")
    if U1 == 1:
        targets = np.zeros((N, 1), np.int64)
    loss = np.empty(N, np.float64)
    grad = np.empty_like(joint) if want_grad else None
    go = _f64(grad_out) if grad_out is not None else None
    rc = lib().ha_oracle_rnnt(_p(joint, _f64p), N, T, U1, V, _p(targets, _i64p), _p(in_len, _i64p),
                              _p(tgt_len, _i64p), int(from_logits), _p(go, _f64p),
                              _p(loss, _f64p), _p(grad, _f64p))
    if rc != 0:
        raise ValueError("ha_oracle_rnnt: bad lengths")
    return loss, grad


def rnnt_fg(f, g, targets, in_len, tgt_len, grad_out=None):
    """Joint-free RNN-T: f (N,T,V), g (N,U+1,V) raw logits -> (loss (N,), grad_f, grad_g).  The joint of
    ha/recognizer.py:114, joint = f[:, :, None, :] + g[:, None, :, :], is built here in float64 and handed to
    rnnt(); the gradients are its reductions over u and over t."""
    f = _f64(f); g = _f64(g)
    joint = np.ascontiguousarray(f[:, :, None, :] + g[:, None, :, :])
    loss, gj = rnnt(joint, targets, in_len, tgt_len, from_logits=True, grad_out=grad_out)
    return loss, gj.sum(axis=2), gj.sum(axis=1)


def greedy(x, in_len=None):
    """x (N,T,V) -> alignment (N,T) i64, score (N,T) f64, hyp (N,T) i64 (-1 padded), hyp_len (N,).
    ha/recognizer.py:48-59 (in_len=None reproduces the reference, which ignores lengths)."""
    x = _f64(x)
    N, T, V = x.shape
    il = _i64(in_len) if in_len is not None else None
    ali = np.empty((N, T), np.int64); sc = np.empty((N, T), np.float64)
    hyp = np.empty((N, T), np.int64); hl = np.empty(N, np.int64)
    lib().ha_oracle_greedy(_p(x, _f64p), N, T, V, _p(il, _i64p), _p(ali, _i64p), _p(sc, _f64p),
                           _p(hyp, _i64p), _p(hl, _i64p))
    return ali, sc, hyp, hl


def ctc_viterbi(lp32, targets, in_len, tgt_len):
    """lp32 (T,N,V) float32 log-probs -> align (N,T) i64 (-1 padded), score (N,) f32.
    Max-semiring variant of ha/ctc.py:144-167; not in the reference (parity unpinned)."""
    lp32 = np.ascontiguousarray(np.asarray(lp32, np.float32))
    targets = _i64(targets); in_len = _i64(in_len); tgt_len = _i64(tgt_len)
    T, N, V = lp32.shape
    S = targets.shape[1]
    ali = np.empty((N, T), np.int64); sc = np.empty(N, np.float32)
    lib().ha_oracle_ctc_viterbi(_p(lp32, _f32p), T, N, V, _p(targets, _i64p), S, _p(in_len, _i64p),
                                _p(tgt_len, _i64p), _p(ali, _i64p), _p(sc, _f32p))
    return ali, sc

"""PyTorch custom ops (with autograd) over the C-ABI library: the thin shim between the
reference-named Python functions and libha_b200.so.

Forward ops return (loss, workspace); the workspace tensor is the saved-for-backward state
(row log-sum-exps and posterior occupancies), never a second B x T x V tensor.  Backward ops
return the gradient with the SAME strides as the input view (ha/recognizer.py:70,78 hands the loss
a permuted view of an (N,T,C) buffer), multiplied by the per-utterance grad_output.
"""
import contextlib

import torch

from . import _lib

_F32 = torch.float32


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_cuda_f32(x, name):
    if not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor: haloop_b200 has no CPU path")
    if x.dtype != _F32:
        raise ValueError(f"{name} must be float32 (got {x.dtype}); the reference forces fp32 too "
                         "(ha/recognizer.py:68-70)")


def _idx(t, device, name):
    """int32/int64 index tensor on `device`, contiguous; returns (tensor, is64)."""
    if t.dtype not in (torch.int32, torch.int64):
        if t.is_floating_point() or t.dtype == torch.bool:
            raise ValueError(f"{name} must be an integer tensor")
        t = t.to(torch.int64)
    if t.device != device:
        t = t.to(device)
    return t.contiguous(), int(t.dtype == torch.int64)


_NULL = contextlib.nullcontext()


def _on(device):
    """Device guard for the C ABI call; the context manager is skipped (it costs more than the call) when the tensor's
    device already is the current one."""
    return _NULL if device.index == torch.cuda.current_device() else torch.cuda.device(device)


def _unit_class_stride(x):
    return x if x.stride(-1) == 1 else x.contiguous()


def _tma_view(x):
    """Unit class stride and, when the class count allows the bulk-copy (TMA) path, rows that start on
    16-byte boundaries; anything else (a sliced view) is copied once."""
    x = _unit_class_stride(x)
    if x.shape[-1] % 4 == 0 and (x.data_ptr() % 16 or any(st % 4 for st in x.stride()[:-1])):
        x = x.contiguous()
    return x


def _empty_like_strided(x):
    """Fresh tensor with x's shape and strides (when x is dense), else contiguous."""
    if x.is_contiguous() or x.numel() == 0:
        return torch.empty_like(x, memory_format=torch.contiguous_format)
    try:
        # dense + non-overlapping views (a permuted contiguous buffer) keep their strides
        perm = sorted(range(x.dim()), key=lambda d: (-x.stride(d), d))
        dense = True
        expect = 1
        for d in reversed(perm):
            if x.size(d) != 1 and x.stride(d) != expect:
                dense = False
                break
            expect *= x.size(d)
        if dense:
            return torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
    except Exception:
        pass
    return torch.empty(x.shape, dtype=x.dtype, device=x.device)


# ------------------------------------------------------------------------------------------- CTC
@torch.library.custom_op("ha_b200::ctc_fwd", mutates_args=())
def ctc_fwd(x: torch.Tensor, targets: torch.Tensor, in_len: torch.Tensor, tgt_len: torch.Tensor,
            from_logits: bool) -> tuple[torch.Tensor, torch.Tensor]:
    _check_cuda_f32(x, "emissions")
    x = _tma_view(x)
    T, N, V = x.shape
    tg, tg64 = _idx(targets, x.device, "targets")
    il, il64 = _idx(in_len, x.device, "emission_lengths")
    tl, tl64 = _idx(tgt_len, x.device, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    if tg.dim() != 2 or tg.shape[0] != N or il.shape != (N,) or tl.shape != (N,):
        raise ValueError("expected targets (N,S), emission_lengths (N,), target_lengths (N,)")
    S = tg.shape[1]
    L = _lib.lib()
    nbytes = L.ha_ctc_workspace_bytes(T, N, V, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    loss = torch.empty(N, dtype=_F32, device=x.device)
    with _on(x.device):
        rc = L.ha_ctc_fwd(x.data_ptr(), x.stride(0), x.stride(1), T, N, V,
                          tg.data_ptr() if S else None, tg.stride(0) if S else 0, S, tg64,
                          il.data_ptr(), tl.data_ptr(), il64, int(from_logits),
                          loss.data_ptr(), ws.data_ptr(), nbytes, _stream(x))
    _lib.check(rc, "ha_ctc_fwd")
    return loss, ws


@ctc_fwd.register_fake
def _(x, targets, in_len, tgt_len, from_logits):
    T, N, V = x.shape
    nbytes = _lib.lib().ha_ctc_workspace_bytes(int(T), int(N), int(V), int(targets.shape[1]))   # host-only query
    return x.new_empty(N), x.new_empty(nbytes, dtype=torch.uint8)


@torch.library.custom_op("ha_b200::ctc_bwd", mutates_args=())
def ctc_bwd(x: torch.Tensor, ws: torch.Tensor, grad_loss: torch.Tensor, S: int,
            from_logits: bool) -> torch.Tensor:
    xs = _tma_view(x)
    T, N, V = xs.shape
    gx = _empty_like_strided(xs)
    g = grad_loss.to(_F32).contiguous()
    L = _lib.lib()
    with _on(x.device):
        rc = L.ha_ctc_bwd(xs.data_ptr(), xs.stride(0), xs.stride(1), T, N, V, S, g.data_ptr(),
                          int(from_logits), gx.data_ptr(), gx.stride(0), gx.stride(1),
                          ws.data_ptr(), ws.numel(), _stream(x))
    _lib.check(rc, "ha_ctc_bwd")
    return gx


@ctc_bwd.register_fake
def _(x, ws, grad_loss, S, from_logits):
    return torch.empty_like(x)


def _ctc_setup(ctx, inputs, output):
    x, targets, _, _, from_logits = inputs
    _, ws = output
    ctx.save_for_backward(x, ws)
    ctx.S = targets.shape[1]
    ctx.from_logits = from_logits


def _ctc_backward(ctx, grad_loss, _grad_ws):
    x, ws = ctx.saved_tensors
    return ctc_bwd(x, ws, grad_loss, ctx.S, ctx.from_logits), None, None, None, None


ctc_fwd.register_autograd(_ctc_backward, setup_context=_ctc_setup)


# -------------------------------------------------------------------------------------- star-CTC
@torch.library.custom_op("ha_b200::star_fwd", mutates_args=())
def star_fwd(x: torch.Tensor, targets: torch.Tensor, in_len: torch.Tensor, tgt_len: torch.Tensor,
             star_penalty: float, from_logits: bool) -> tuple[torch.Tensor, torch.Tensor]:
    _check_cuda_f32(x, "emissions")
    x = _tma_view(x)
    T, N, V = x.shape
    tg, tg64 = _idx(targets, x.device, "targets")
    il, il64 = _idx(in_len, x.device, "emission_lengths")
    tl, tl64 = _idx(tgt_len, x.device, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    if tg.dim() != 2 or tg.shape[0] != N or il.shape != (N,) or tl.shape != (N,):
        raise ValueError("expected targets (N,S), emission_lengths (N,), target_lengths (N,)")
    S = tg.shape[1]
    L = _lib.lib()
    nbytes = L.ha_star_workspace_bytes(T, N, V, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    loss = torch.empty(N, dtype=_F32, device=x.device)
    with _on(x.device):
        rc = L.ha_star_fwd(x.data_ptr(), x.stride(0), x.stride(1), T, N, V,
                           tg.data_ptr() if S else None, tg.stride(0) if S else 0, S, tg64,
                           il.data_ptr(), tl.data_ptr(), il64, float(star_penalty), int(from_logits),
                           loss.data_ptr(), ws.data_ptr(), nbytes, _stream(x))
    _lib.check(rc, "ha_star_fwd")
    return loss, ws


@star_fwd.register_fake
def _(x, targets, in_len, tgt_len, star_penalty, from_logits):
    T, N, V = x.shape
    nbytes = _lib.lib().ha_star_workspace_bytes(int(T), int(N), int(V), int(targets.shape[1]))
    return x.new_empty(N), x.new_empty(nbytes, dtype=torch.uint8)


@torch.library.custom_op("ha_b200::star_bwd", mutates_args=())
def star_bwd(x: torch.Tensor, ws: torch.Tensor, grad_loss: torch.Tensor, S: int,
             from_logits: bool) -> torch.Tensor:
    xs = _tma_view(x)
    T, N, V = xs.shape
    gx = _empty_like_strided(xs)
    g = grad_loss.to(_F32).contiguous()
    L = _lib.lib()
    with _on(x.device):
        rc = L.ha_star_bwd(xs.data_ptr(), xs.stride(0), xs.stride(1), T, N, V, S, g.data_ptr(),
                           int(from_logits), gx.data_ptr(), gx.stride(0), gx.stride(1),
                           ws.data_ptr(), ws.numel(), _stream(x))
    _lib.check(rc, "ha_star_bwd")
    return gx


@star_bwd.register_fake
def _(x, ws, grad_loss, S, from_logits):
    return torch.empty_like(x)


def _star_setup(ctx, inputs, output):
    x, targets, _, _, _, from_logits = inputs
    _, ws = output
    ctx.save_for_backward(x, ws)
    ctx.S = targets.shape[1]
    ctx.from_logits = from_logits


def _star_backward(ctx, grad_loss, _grad_ws):
    x, ws = ctx.saved_tensors
    return star_bwd(x, ws, grad_loss, ctx.S, ctx.from_logits), None, None, None, None, None


star_fwd.register_autograd(_star_backward, setup_context=_star_setup)


# ------------------------------------------------------------------------- no second derivative
def _no_double_backward(op, name):
    """The backward ops return a first-order gradient computed by a kernel; nothing differentiates through it.
    The reference's pure-PyTorch losses support double backward (gradient penalties, create_graph=True): make
    that use fail loudly here instead of silently treating the gradient as a constant."""
    def _raise(ctx, *grads):
        raise NotImplementedError(f"haloop_b200: {name} has no second derivative (double backward / "
                                  "create_graph=True through the alignment losses is not supported)")
    op.register_autograd(_raise)


# ------------------------------------------------------------------------------------------ vmap
# Utterances are independent, so the batching rule of every op is "fold the vmapped dimension into the
# utterance dimension and run once" (the reference has to drop its CTC term under torch.func.vmap for lack
# of a rule, ha/grad_norm.py:14-16).  The workspace of the merged call is returned unbatched: the matching
# backward op is vmapped the same way and finds the merged layout.
def _fold(t, d, B, at):
    """Move vmap dim `d` of `t` (or broadcast if None) next to dim `at` and merge the two."""
    if t is None:
        return None
    if d is None:
        t = t.unsqueeze(at).expand(*t.shape[:at], B, *t.shape[at:])
    else:
        t = t.movedim(d, at)
    return t.reshape(*t.shape[:at], t.shape[at] * t.shape[at + 1], *t.shape[at + 2:])


def _bsize(args, in_dims):
    for a, d in zip(args, in_dims):
        if d is not None:
            return a.shape[d]
    raise RuntimeError("vmap rule called without a batched argument")


def _register_time_major_vmap(fwd, bwd, n_extra):
    """fwd(x (T,N,V), targets (N,S), in_len (N,), tgt_len (N,), *extra) / bwd(x, ws, grad_loss (N,), S, flag)"""
    @torch.library.register_vmap(fwd)
    def _(info, in_dims, x, targets, in_len, tgt_len, *extra):
        B = _bsize((x, targets, in_len, tgt_len), in_dims[:4])
        loss, ws = fwd(_fold(x, in_dims[0], B, 1), _fold(targets, in_dims[1], B, 0), _fold(in_len, in_dims[2], B, 0),
                       _fold(tgt_len, in_dims[3], B, 0), *extra)
        return (loss.view(B, -1), ws), (0, None)

    @torch.library.register_vmap(bwd)
    def _(info, in_dims, x, ws, grad_loss, S, flag):
        if in_dims[1] is not None:
            raise RuntimeError("the workspace of a vmapped forward is shared by the whole batch")
        B = _bsize((x, grad_loss), (in_dims[0], in_dims[2]))
        gx = bwd(_fold(x, in_dims[0], B, 1), ws, _fold(grad_loss, in_dims[2], B, 0), S, flag)
        T, BN, V = gx.shape
        return gx.view(T, B, BN // B, V), 1


# ----------------------------------------------------------------------------------------- RNN-T
@torch.library.custom_op("ha_b200::rnnt_fwd", mutates_args=())
def rnnt_fwd(joint: torch.Tensor, targets: torch.Tensor, in_len: torch.Tensor, tgt_len: torch.Tensor,
             from_logits: bool) -> tuple[torch.Tensor, torch.Tensor]:
    _check_cuda_f32(joint, "joint")
    joint = _unit_class_stride(joint)            # any (n, t, u) strides are taken as they are: no copy of the joint
    N, T, U1, V = joint.shape
    tg, tg64 = _idx(targets, joint.device, "targets")
    il, il64 = _idx(in_len, joint.device, "joint_lengths")
    tl, tl64 = _idx(tgt_len, joint.device, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    if tg.dim() != 2 or tg.shape != (N, U1 - 1) or il.shape != (N,) or tl.shape != (N,):
        raise ValueError("expected joint (N,T,U+1,V), targets (N,U), joint_lengths (N,), target_lengths (N,)")
    L = _lib.lib()
    nbytes = L.ha_rnnt_workspace_bytes(N, T, U1, V)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=joint.device)
    loss = torch.empty(N, dtype=_F32, device=joint.device)
    with _on(joint.device):
        rc = L.ha_rnnt_fwd(joint.data_ptr(), joint.stride(0), joint.stride(1), joint.stride(2), N, T, U1, V,
                           tg.data_ptr() if U1 > 1 else None, tg.stride(0) if U1 > 1 else 0, tg64,
                           il.data_ptr(), tl.data_ptr(), il64, int(from_logits),
                           loss.data_ptr(), ws.data_ptr(), nbytes, _stream(joint))
    _lib.check(rc, "ha_rnnt_fwd")
    return loss, ws


@rnnt_fwd.register_fake
def _(joint, targets, in_len, tgt_len, from_logits):
    N, T, U1, V = joint.shape
    nbytes = _lib.lib().ha_rnnt_workspace_bytes(int(N), int(T), int(U1), int(V))
    return joint.new_empty(N), joint.new_empty(nbytes, dtype=torch.uint8)


@torch.library.custom_op("ha_b200::rnnt_bwd", mutates_args=())
def rnnt_bwd(joint: torch.Tensor, ws: torch.Tensor, grad_loss: torch.Tensor,
             from_logits: bool) -> torch.Tensor:
    joint = _unit_class_stride(joint)
    N, T, U1, V = joint.shape
    gj = _empty_like_strided(joint)              # the gradient takes the strides of the joint view
    g = grad_loss.to(_F32).contiguous()
    L = _lib.lib()
    with _on(joint.device):
        rc = L.ha_rnnt_bwd(joint.data_ptr(), joint.stride(0), joint.stride(1), joint.stride(2), N, T, U1, V,
                           g.data_ptr(), int(from_logits), gj.data_ptr(), gj.stride(0), gj.stride(1), gj.stride(2),
                           ws.data_ptr(), ws.numel(), _stream(joint))
    _lib.check(rc, "ha_rnnt_bwd")
    return gj


@rnnt_bwd.register_fake
def _(joint, ws, grad_loss, from_logits):
    return torch.empty_like(joint)


def _rnnt_setup(ctx, inputs, output):
    joint, _, _, _, from_logits = inputs
    _, ws = output
    ctx.save_for_backward(joint, ws)
    ctx.from_logits = from_logits


def _rnnt_backward(ctx, grad_loss, _grad_ws):
    joint, ws = ctx.saved_tensors
    return rnnt_bwd(joint, ws, grad_loss, ctx.from_logits), None, None, None, None


rnnt_fwd.register_autograd(_rnnt_backward, setup_context=_rnnt_setup)


# --------------------------------------------------------------------------- joint-free RNN-T
@torch.library.custom_op("ha_b200::rnnt_fg_fwd", mutates_args=())
def rnnt_fg_fwd(f: torch.Tensor, g: torch.Tensor, targets: torch.Tensor, in_len: torch.Tensor,
                tgt_len: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    _check_cuda_f32(f, "f")
    _check_cuda_f32(g, "g")
    f = f.contiguous(); g = g.contiguous()
    if f.dim() != 3 or g.dim() != 3 or f.shape[0] != g.shape[0] or f.shape[2] != g.shape[2]:
        raise ValueError("expected f (N,T,V) and g (N,U+1,V)")
    N, T, V = f.shape
    U1 = g.shape[1]
    tg, tg64 = _idx(targets, f.device, "targets")
    il, il64 = _idx(in_len, f.device, "joint_lengths")
    tl, tl64 = _idx(tgt_len, f.device, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    if tg.dim() != 2 or tg.shape != (N, U1 - 1) or il.shape != (N,) or tl.shape != (N,):
        raise ValueError("expected f (N,T,V), g (N,U+1,V), targets (N,U), joint_lengths (N,), target_lengths (N,)")
    L = _lib.lib()
    nbytes = L.ha_rnnt_fg_workspace_bytes(N, T, U1, V)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=f.device)
    loss = torch.empty(N, dtype=_F32, device=f.device)
    with _on(f.device):
        rc = L.ha_rnnt_fg_fwd(f.data_ptr(), g.data_ptr(), N, T, U1, V,
                              tg.data_ptr() if U1 > 1 else None, tg.stride(0) if U1 > 1 else 0, tg64,
                              il.data_ptr(), tl.data_ptr(), il64,
                              loss.data_ptr(), ws.data_ptr(), nbytes, _stream(f))
    _lib.check(rc, "ha_rnnt_fg_fwd")
    return loss, ws


@rnnt_fg_fwd.register_fake
def _(f, g, targets, in_len, tgt_len):
    N, T, V = f.shape
    nbytes = _lib.lib().ha_rnnt_fg_workspace_bytes(int(N), int(T), int(g.shape[1]), int(V))
    return f.new_empty(N), f.new_empty(nbytes, dtype=torch.uint8)


@torch.library.custom_op("ha_b200::rnnt_fg_bwd", mutates_args=())
def rnnt_fg_bwd(f: torch.Tensor, g: torch.Tensor, ws: torch.Tensor,
                grad_loss: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    f = f.contiguous(); g = g.contiguous()
    N, T, V = f.shape
    U1 = g.shape[1]
    gf = torch.empty_like(f, memory_format=torch.contiguous_format)
    gg = torch.empty_like(g, memory_format=torch.contiguous_format)
    go = grad_loss.to(_F32).contiguous()
    L = _lib.lib()
    with _on(f.device):
        rc = L.ha_rnnt_fg_bwd(f.data_ptr(), g.data_ptr(), N, T, U1, V, go.data_ptr(), gf.data_ptr(), gg.data_ptr(),
                              ws.data_ptr(), ws.numel(), _stream(f))
    _lib.check(rc, "ha_rnnt_fg_bwd")
    return gf, gg


@rnnt_fg_bwd.register_fake
def _(f, g, ws, grad_loss):
    return torch.empty_like(f), torch.empty_like(g)


def _rnnt_fg_setup(ctx, inputs, output):
    f, g, _, _, _ = inputs
    _, ws = output
    ctx.save_for_backward(f, g, ws)


def _rnnt_fg_backward(ctx, grad_loss, _grad_ws):
    f, g, ws = ctx.saved_tensors
    gf, gg = rnnt_fg_bwd(f, g, ws, grad_loss)
    return gf, gg, None, None, None


rnnt_fg_fwd.register_autograd(_rnnt_fg_backward, setup_context=_rnnt_fg_setup)


# ------------------------------------------------------------------------------------- alignment
def greedy_decode(x, in_len=None):
    """x (N,T,V) -> alignment (N,T) i64, score (N,T) f32, hyp (N,T) i64 padded with -1, hyp_len (N,)."""
    _check_cuda_f32(x, "log_probs")
    x = _unit_class_stride(x)
    N, T, V = x.shape
    dev = x.device
    ali = torch.empty(N, T, dtype=torch.int64, device=dev)
    sc = torch.empty(N, T, dtype=_F32, device=dev)
    hyp = torch.empty(N, T, dtype=torch.int64, device=dev)
    hl = torch.empty(N, dtype=torch.int64, device=dev)
    il, il64 = (None, 0) if in_len is None else _idx(in_len, dev, "input_lengths")
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.ha_greedy_decode(x.data_ptr(), x.stride(0), x.stride(1), N, T, V,
                                il.data_ptr() if il is not None else None, il64,
                                ali.data_ptr(), sc.data_ptr(), hyp.data_ptr(), hl.data_ptr(), _stream(x))
    _lib.check(rc, "ha_greedy_decode")
    return ali, sc, hyp, hl


# --------------------------------------------------------------------- fused classifier head + CTC
_PRECISION = {"tf32x3": 3, "tf32": 1}


def _head_sizes(N, T, D, V, S):
    import ctypes
    a, b, c = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = _lib.lib().ha_head_ctc_workspace_bytes(N, T, D, V, S, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    _lib.check(rc, "ha_head_ctc_workspace_bytes")
    return a.value, b.value, c.value


@torch.library.custom_op("ha_b200::head_ctc_fwd", mutates_args=())
def head_ctc_fwd(h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, targets: torch.Tensor,
                 in_len: torch.Tensor, tgt_len: torch.Tensor, precision: int) -> tuple[torch.Tensor, torch.Tensor]:
    """h (N,T,D), weight (V,D), bias (V) -> per-utterance CTC loss (N,) of log_softmax(h W^T + b), saved state."""
    _check_cuda_f32(h, "features")
    _check_cuda_f32(weight, "weight")
    h, weight = h.contiguous(), weight.contiguous()
    N, T, D = h.shape
    V = weight.shape[0]
    if weight.shape[1] != D:
        raise ValueError("weight must be (V, D)")
    b = None if bias is None else bias.to(_F32).contiguous()
    tg, tg64 = _idx(targets, h.device, "targets")
    il, il64 = _idx(in_len, h.device, "input_lengths")
    tl, tl64 = _idx(tgt_len, h.device, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    if tg.dim() != 2 or tg.shape[0] != N or il.shape != (N,) or tl.shape != (N,):
        raise ValueError("expected targets (N,S), input_lengths (N,), target_lengths (N,)")
    S = tg.shape[1]
    saved_b, fwd_b, _ = _head_sizes(N, T, D, V, S)
    saved = torch.empty(saved_b, dtype=torch.uint8, device=h.device)
    scratch = torch.empty(fwd_b, dtype=torch.uint8, device=h.device)
    loss = torch.empty(N, dtype=_F32, device=h.device)
    with _on(h.device):
        rc = _lib.lib().ha_head_ctc_fwd(h.data_ptr(), weight.data_ptr(), b.data_ptr() if b is not None else None,
                                        N, T, D, V, tg.data_ptr() if S else None, tg.stride(0) if S else 0, S, tg64,
                                        il.data_ptr(), tl.data_ptr(), il64, int(precision), loss.data_ptr(),
                                        saved.data_ptr(), saved_b, scratch.data_ptr(), fwd_b, _stream(h))
    _lib.check(rc, "ha_head_ctc_fwd")
    return loss, saved


@head_ctc_fwd.register_fake
def _(h, weight, bias, targets, in_len, tgt_len, precision):
    N, T, D = h.shape
    saved_b, _, _ = _head_sizes(int(N), int(T), int(D), int(weight.shape[0]), int(targets.shape[1]))
    return h.new_empty(N), h.new_empty(saved_b, dtype=torch.uint8)


@torch.library.custom_op("ha_b200::head_ctc_bwd", mutates_args=())
def head_ctc_bwd(h: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, saved: torch.Tensor,
                 grad_loss: torch.Tensor, S: int, precision: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    h, weight = h.contiguous(), weight.contiguous()
    N, T, D = h.shape
    V = weight.shape[0]
    b = None if bias is None else bias.to(_F32).contiguous()
    g = grad_loss.to(_F32).contiguous()
    _, _, bwd_b = _head_sizes(N, T, D, V, S)
    scratch = torch.empty(bwd_b, dtype=torch.uint8, device=h.device)
    dh = torch.empty_like(h)
    dW = torch.empty_like(weight)
    db = torch.empty(V, dtype=_F32, device=h.device)
    with _on(h.device):
        rc = _lib.lib().ha_head_ctc_bwd(h.data_ptr(), weight.data_ptr(), b.data_ptr() if b is not None else None,
                                        N, T, D, V, S, g.data_ptr(), int(precision), dh.data_ptr(), dW.data_ptr(),
                                        db.data_ptr(), saved.data_ptr(), saved.numel(), scratch.data_ptr(), bwd_b,
                                        _stream(h))
    _lib.check(rc, "ha_head_ctc_bwd")
    return dh, dW, db


@head_ctc_bwd.register_fake
def _(h, weight, bias, saved, grad_loss, S, precision):
    return torch.empty_like(h), torch.empty_like(weight), weight.new_empty(weight.shape[0])


def _head_setup(ctx, inputs, output):
    h, weight, bias, targets, _, _, precision = inputs
    _, saved = output
    ctx.save_for_backward(h, weight, bias, saved)
    ctx.S = targets.shape[1]
    ctx.precision = precision
    ctx.has_bias = bias is not None


def _head_backward(ctx, grad_loss, _grad_saved):
    h, weight, bias, saved = ctx.saved_tensors
    if grad_loss is None:
        grad_loss = torch.zeros(h.shape[0], dtype=_F32, device=h.device)
    dh, dW, db = head_ctc_bwd(h, weight, bias, saved, grad_loss, ctx.S, ctx.precision)
    return dh, dW, (db if ctx.has_bias else None), None, None, None, None


head_ctc_fwd.register_autograd(_head_backward, setup_context=_head_setup)


def ctc_beam_search(x, in_len=None, beam_size=3, reference_ext_blank=True):
    """x (N,T,V) log-probs -> hyp (N,beam,T) i64 padded with -1 (best first), hyp_len (N,beam), score (N,beam)."""
    _check_cuda_f32(x, "log_probs")
    x = _unit_class_stride(x)
    N, T, V = x.shape
    dev = x.device
    hyp = torch.empty(N, beam_size, T, dtype=torch.int64, device=dev)
    hl = torch.empty(N, beam_size, dtype=torch.int64, device=dev)
    sc = torch.empty(N, beam_size, dtype=_F32, device=dev)
    il, il64 = (None, 0) if in_len is None else _idx(in_len, dev, "input_lengths")
    L = _lib.lib()
    nbytes = L.ha_ctc_beam_search_workspace_bytes(N, T, V, beam_size)
    if nbytes == 0:
        raise ValueError("beam_size must be between 1 and 16")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.ha_ctc_beam_search(x.data_ptr(), x.stride(0), x.stride(1), N, T, V,
                                  il.data_ptr() if il is not None else None, il64, int(beam_size), int(bool(reference_ext_blank)),
                                  hyp.data_ptr(), hl.data_ptr(), sc.data_ptr(), ws.data_ptr(), nbytes, _stream(x))
    _lib.check(rc, "ha_ctc_beam_search")
    return hyp, hl, sc


def ctc_viterbi(lp, targets, in_len, tgt_len):
    """lp (T,N,V) log-probs -> alignment (N,T) i64 (class per frame, -1 padded), score (N,) f32."""
    _check_cuda_f32(lp, "log_probs")
    lp = _unit_class_stride(lp)
    T, N, V = lp.shape
    dev = lp.device
    tg, tg64 = _idx(targets, dev, "targets")
    il, il64 = _idx(in_len, dev, "input_lengths")
    tl, tl64 = _idx(tgt_len, dev, "target_lengths")
    if il64 != tl64:
        il, tl, il64 = il.to(torch.int64), tl.to(torch.int64), 1
    S = tg.shape[1]
    L = _lib.lib()
    nbytes = L.ha_ctc_viterbi_workspace_bytes(T, N, V, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ali = torch.empty(N, T, dtype=torch.int64, device=dev)
    sc = torch.empty(N, dtype=_F32, device=dev)
    with torch.cuda.device(dev):
        rc = L.ha_ctc_viterbi(lp.data_ptr(), lp.stride(0), lp.stride(1), T, N, V,
                              tg.data_ptr() if S else None, tg.stride(0) if S else 0, S, tg64,
                              il.data_ptr(), tl.data_ptr(), il64,
                              ali.data_ptr(), sc.data_ptr(), ws.data_ptr(), nbytes, _stream(lp))
    _lib.check(rc, "ha_ctc_viterbi")
    return ali, sc


_no_double_backward(ctc_bwd, "ctc_bwd")
_no_double_backward(star_bwd, "star_bwd")
_no_double_backward(rnnt_bwd, "rnnt_bwd")
_no_double_backward(rnnt_fg_bwd, "rnnt_fg_bwd")
_no_double_backward(head_ctc_bwd, "head_ctc_bwd")

# ------------------------------------------------------------------------------- vmap rules
_register_time_major_vmap(ctc_fwd, ctc_bwd, 1)
_register_time_major_vmap(star_fwd, star_bwd, 2)


@torch.library.register_vmap(rnnt_fwd)
def _(info, in_dims, joint, targets, in_len, tgt_len, from_logits):
    B = _bsize((joint, targets, in_len, tgt_len), in_dims[:4])
    j = _fold(joint, in_dims[0], B, 0)
    loss, ws = rnnt_fwd(j, _fold(targets, in_dims[1], B, 0), _fold(in_len, in_dims[2], B, 0),
                        _fold(tgt_len, in_dims[3], B, 0), from_logits)
    return (loss.view(B, -1), ws), (0, None)


@torch.library.register_vmap(rnnt_bwd)
def _(info, in_dims, joint, ws, grad_loss, from_logits):
    if in_dims[1] is not None:
        raise RuntimeError("the workspace of a vmapped forward is shared by the whole batch")
    B = _bsize((joint, grad_loss), (in_dims[0], in_dims[2]))
    gj = rnnt_bwd(_fold(joint, in_dims[0], B, 0), ws, _fold(grad_loss, in_dims[2], B, 0), from_logits)
    return gj.view(B, gj.shape[0] // B, *gj.shape[1:]), 0


@torch.library.register_vmap(rnnt_fg_fwd)
def _(info, in_dims, f, g, targets, in_len, tgt_len):
    B = _bsize((f, g, targets, in_len, tgt_len), in_dims)
    loss, ws = rnnt_fg_fwd(_fold(f, in_dims[0], B, 0), _fold(g, in_dims[1], B, 0), _fold(targets, in_dims[2], B, 0),
                           _fold(in_len, in_dims[3], B, 0), _fold(tgt_len, in_dims[4], B, 0))
    return (loss.view(B, -1), ws), (0, None)


@torch.library.register_vmap(rnnt_fg_bwd)
def _(info, in_dims, f, g, ws, grad_loss):
    if in_dims[2] is not None:
        raise RuntimeError("the workspace of a vmapped forward is shared by the whole batch")
    B = _bsize((f, g, grad_loss), (in_dims[0], in_dims[1], in_dims[3]))
    gf, gg = rnnt_fg_bwd(_fold(f, in_dims[0], B, 0), _fold(g, in_dims[1], B, 0), ws, _fold(grad_loss, in_dims[3], B, 0))
    return (gf.view(B, gf.shape[0] // B, *gf.shape[1:]), gg.view(B, gg.shape[0] // B, *gg.shape[1:])), (0, 0)

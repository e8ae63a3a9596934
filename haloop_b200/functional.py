"""torch.func (functorch) entry points.

The custom ops in ops.py carry their autograd formula through torch.library.register_autograd, which
torch.func.grad / vjp cannot enter (the generated autograd.Function has no setup_context).  When function
transforms are active, the public wrappers route through these autograd.Functions instead: same kernels,
plus a vmap rule that folds the vmapped dimension into the utterance dimension (utterances are independent),
so vmap(grad_and_value(loss)) -- the per-sample gradient pattern of ha/grad_norm.py, where the reference has
to drop its CTC term for lack of a batching rule (:14-16) -- runs as ONE merged launch.
"""
import torch

from . import ops


def transforms_active():
    try:
        return torch._C._are_functorch_transforms_active()
    except Exception:
        return False


def _fold(t, d, B, at):
    return ops._fold(t, d, B, at)


def _plain(t):
    """Strip torch.func wrappers (grad-tracking levels): the backward kernels are first-order only, and a plain
    tensor is a constant at every active level, so the op then runs as a kernel with no autograd formula."""
    f = torch._C._functorch
    while isinstance(t, torch.Tensor) and f.is_functorch_wrapped_tensor(t) and not f.is_batchedtensor(t):
        t = f.get_unwrapped(t)
    return t


def _below_autograd():
    """Inside these Functions the ops are plain kernels: skip their own (register_autograd) formula, whose
    generated autograd.Function refuses to run while function transforms are active."""
    return torch._C._AutoDispatchBelowAutograd()


class _TimeMajor(torch.autograd.Function):
    """CTC / star-CTC on emissions (T,N,V).  kind: 0 = CTC, 1 = star."""
    generate_vmap_rule = False

    @staticmethod
    def forward(x, targets, in_len, tgt_len, penalty, from_logits, kind):
        with _below_autograd():
            if kind == 0:
                return ops.ctc_fwd(x, targets, in_len, tgt_len, from_logits)
            return ops.star_fwd(x, targets, in_len, tgt_len, penalty, from_logits)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, targets, _, _, _, from_logits, kind = inputs
        ctx.save_for_backward(x, output[1])
        ctx.S, ctx.from_logits, ctx.kind = targets.shape[1], from_logits, kind
        ctx.mark_non_differentiable(output[1])

    @staticmethod
    @torch.autograd.function.once_differentiable     # no second derivative: double backward raises
    def backward(ctx, grad_loss, _grad_ws):
        x, ws = ctx.saved_tensors
        bwd = ops.ctc_bwd if ctx.kind == 0 else ops.star_bwd
        return bwd(_plain(x), _plain(ws), _plain(grad_loss), ctx.S, ctx.from_logits), None, None, None, None, None, None

    @staticmethod
    def vmap(info, in_dims, x, targets, in_len, tgt_len, penalty, from_logits, kind):
        B = info.batch_size
        loss, ws = _TimeMajor.apply(_fold(x, in_dims[0], B, 1), _fold(targets, in_dims[1], B, 0),
                                    _fold(in_len, in_dims[2], B, 0), _fold(tgt_len, in_dims[3], B, 0),
                                    penalty, from_logits, kind)
        return (loss.view(B, -1), ws), (0, None)


class _Joint(torch.autograd.Function):
    generate_vmap_rule = False

    @staticmethod
    def forward(joint, targets, in_len, tgt_len, from_logits):
        with _below_autograd():
            return ops.rnnt_fwd(joint, targets, in_len, tgt_len, from_logits)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], output[1])
        ctx.from_logits = inputs[4]
        ctx.mark_non_differentiable(output[1])

    @staticmethod
    @torch.autograd.function.once_differentiable     # no second derivative: double backward raises
    def backward(ctx, grad_loss, _grad_ws):
        joint, ws = ctx.saved_tensors
        return ops.rnnt_bwd(_plain(joint), _plain(ws), _plain(grad_loss), ctx.from_logits), None, None, None, None

    @staticmethod
    def vmap(info, in_dims, joint, targets, in_len, tgt_len, from_logits):
        B = info.batch_size
        loss, ws = _Joint.apply(_fold(joint, in_dims[0], B, 0), _fold(targets, in_dims[1], B, 0),
                                _fold(in_len, in_dims[2], B, 0), _fold(tgt_len, in_dims[3], B, 0), from_logits)
        return (loss.view(B, -1), ws), (0, None)


class _Factored(torch.autograd.Function):
    generate_vmap_rule = False

    @staticmethod
    def forward(f, g, targets, in_len, tgt_len):
        with _below_autograd():
            return ops.rnnt_fg_fwd(f, g, targets, in_len, tgt_len)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1], output[1])
        ctx.mark_non_differentiable(output[1])

    @staticmethod
    @torch.autograd.function.once_differentiable     # no second derivative: double backward raises
    def backward(ctx, grad_loss, _grad_ws):
        f, g, ws = ctx.saved_tensors
        gf, gg = ops.rnnt_fg_bwd(_plain(f), _plain(g), _plain(ws), _plain(grad_loss))
        return gf, gg, None, None, None

    @staticmethod
    def vmap(info, in_dims, f, g, targets, in_len, tgt_len):
        B = info.batch_size
        loss, ws = _Factored.apply(_fold(f, in_dims[0], B, 0), _fold(g, in_dims[1], B, 0),
                                   _fold(targets, in_dims[2], B, 0), _fold(in_len, in_dims[3], B, 0),
                                   _fold(tgt_len, in_dims[4], B, 0))
        return (loss.view(B, -1), ws), (0, None)


def ctc(x, targets, in_len, tgt_len, from_logits):
    return _TimeMajor.apply(x, targets, in_len, tgt_len, 0.0, bool(from_logits), 0)[0]


def star(x, targets, in_len, tgt_len, penalty, from_logits):
    return _TimeMajor.apply(x, targets, in_len, tgt_len, float(penalty), bool(from_logits), 1)[0]


def rnnt(joint, targets, in_len, tgt_len, from_logits):
    return _Joint.apply(joint, targets, in_len, tgt_len, bool(from_logits))[0]


def rnnt_fg(f, g, targets, in_len, tgt_len):
    return _Factored.apply(f, g, targets, in_len, tgt_len)[0]

"""Fused classifier head + CTC: `TemporalClassifier.log_probs` followed by the CTC loss
(ha/recognizer.py:43-46, 61-82) as ONE op on the encoder features.  The (N,T,V) logits, the log-probs and their
gradient are never materialised; the weight and feature gradients come out of the same tensor-core kernels."""
import torch

from . import ops


def linear_ctc_forward_score(features, weight, bias, targets, input_lengths, target_lengths, precision="tf32x3"):
    """Per-utterance -log p(targets | log_softmax(features @ weight.T + bias)), (N,), differentiable w.r.t.
    features (N,T,D), weight (V,D) and bias (V).  Equivalent to (ha/recognizer.py:43-46 + ha/ctc.py:110-174)

        lp = torch.nn.functional.linear(features, weight, bias).log_softmax(-1)
        ctc_forward_score3(lp.permute(1, 0, 2), targets, input_lengths, target_lengths)

    precision "tf32x3" (default) keeps fp32-grade products on the tf32 tensor cores with a three-product split;
    "tf32" is the single-product mode (what torch does with allow_tf32 = True)."""
    if precision not in ops._PRECISION:
        raise ValueError(f"precision must be one of {sorted(ops._PRECISION)}")
    loss, _ = ops.head_ctc_fwd(features, weight, bias, targets, input_lengths, target_lengths, ops._PRECISION[precision])
    return loss


def linear_ctc_loss(features, weight, bias, targets, input_lengths, target_lengths, reduction="mean",
                    zero_infinity=False, precision="tf32x3"):
    """F.ctc_loss(F.linear(features, weight, bias).log_softmax(-1).permute(1,0,2), ...) with F.ctc_loss's reductions
    (ha/recognizer.py:71 uses the default 'mean': per-utterance loss / target length (clamped at 1), batch mean)."""
    losses = linear_ctc_forward_score(features, weight, bias, targets, input_lengths, target_lengths, precision)
    if zero_infinity:
        losses = torch.where(torch.isinf(losses), torch.zeros_like(losses), losses)
    if reduction == "none":
        return losses
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean":
        tl = target_lengths.to(losses.device).clamp_min(1).to(losses.dtype)
        return (losses / tl).mean()
    raise ValueError(f"unknown reduction {reduction!r}")

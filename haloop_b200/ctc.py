"""Drop-in for ha/ctc.py: same names, same positional signatures, CUDA hot path.

    ctc_forward_score3(emissions, targets, emission_lengths, target_lengths) -> (N,)   ha/ctc.py:110-174
    ctc_reduce_mean(losses, target_lengths) -> scalar                                  ha/ctc.py:177-178
"""
import torch

from . import functional, ops


def ctc_forward_score3(emissions, targets, emission_lengths, target_lengths, from_logits=False):
    """CTC negative log-likelihood per utterance.

    emissions (T,N,C) float32 CUDA, any t/n strides with unit class stride (the permuted view of
    ha/recognizer.py:70 is consumed as is); targets (N,S) 0-padded; blank = 0.
    from_logits=False is the reference contract: `emissions` are log-probs, used as given, and
    autograd returns -occupancy * grad at this boundary.  from_logits=True takes raw logits, fuses
    the log-softmax and returns (softmax - occupancy) * grad: one read and one write of (T,N,C).
    Infeasible alignments give +inf and a zero gradient (the reference gives ~3.4e38).
    """
    if functional.transforms_active():          # torch.func.grad / vmap: see functional.py
        return functional.ctc(emissions, targets, emission_lengths, target_lengths, from_logits)
    loss, _ = ops.ctc_fwd(emissions, targets, emission_lengths, target_lengths, bool(from_logits))
    return loss


def ctc_reduce_mean(losses, target_lengths):
    """(losses / target_lengths).mean(-1), exactly as ha/ctc.py:177-178: like the reference it divides by the raw
    target length, so an empty transcript gives inf (F.ctc_loss(reduction='mean') clamps the length at 1 instead;
    use ctc_loss(..., reduction='mean') or the patched TemporalClassifier.forward for that behaviour)."""
    return (losses / target_lengths.to(losses.device)).mean(-1)


def ctc_loss(log_probs, targets, input_lengths, target_lengths, reduction="mean", from_logits=False):
    """torch.nn.functional.ctc_loss-compatible wrapper for the live call site ha/recognizer.py:71
    (blank=0, zero_infinity=False)."""
    losses = ctc_forward_score3(log_probs, targets, input_lengths, target_lengths, from_logits=from_logits)
    if reduction == "none":
        return losses
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean":
        tl = target_lengths.to(losses.device).clamp_min(1)
        return (losses / tl).mean()
    raise ValueError(f"unknown reduction {reduction!r}")

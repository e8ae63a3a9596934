"""Drop-in for ha/transducer.py: RNN-T lattice loss.

    transducer_forward_score(joint, targets, joint_lengths, target_lengths) -> (N,)   ha/transducer.py:175-205
"""
from . import ops


def transducer_forward_score(joint, targets, joint_lengths, target_lengths, from_logits=False):
    """RNN-T negative log-likelihood per utterance.

    joint (N,T,U+1,K) float32 CUDA: log-softmaxed (reference contract, from_logits=False) or raw
    joint logits (from_logits=True: the log-softmax is fused and the (N,T,U+1,K) tensor is read twice
    and written once in total).  targets (N,U); blank = 0.  Unlike the reference there is no
    power-of-two restriction on T (ha/transducer.py:194-195) and no -10000 scan seed (ha/scan.py:116).
    """
    loss, _ = ops.rnnt_fwd(joint, targets, joint_lengths, target_lengths, bool(from_logits))
    return loss


def rnnt_loss(logits, targets, logit_lengths, target_lengths, blank=0, reduction="mean",
              fused_log_softmax=True):
    """torchaudio.functional.rnnt_loss-compatible wrapper for the live call site
    ha/recognizer.py:121-126 ('mean' is a plain batch mean, as in torchaudio)."""
    if blank != 0:
        raise ValueError("blank must be 0 (ha/transducer.py:190)")
    losses = transducer_forward_score(logits, targets, logit_lengths, target_lengths,
                                      from_logits=bool(fused_log_softmax))
    if reduction == "none":
        return losses
    if reduction == "sum":
        return losses.sum()
    if reduction == "mean":
        return losses.mean()
    raise ValueError(f"unknown reduction {reduction!r}")

"""Drop-in for ha/transducer.py: RNN-T lattice loss.

    transducer_forward_score(joint, targets, joint_lengths, target_lengths) -> (N,)   ha/transducer.py:175-205
    transducer_forward_score_fg(f, g, targets, joint_lengths, target_lengths) -> (N,)  joint-free variant
"""
from . import functional, ops


def transducer_forward_score(joint, targets, joint_lengths, target_lengths, from_logits=False):
    """RNN-T negative log-likelihood per utterance.

    joint (N,T,U+1,K) float32 CUDA: log-softmaxed (reference contract, from_logits=False) or raw
    joint logits (from_logits=True: the log-softmax is fused and the (N,T,U+1,K) tensor is read twice
    and written once in total).  targets (N,U); blank = 0.  Unlike the reference there is no
    power-of-two restriction on T (ha/transducer.py:194-195) and no -10000 scan seed (ha/scan.py:116).
    """
    if functional.transforms_active():          # torch.func.grad / vmap: see functional.py
        return functional.rnnt(joint, targets, joint_lengths, target_lengths, from_logits)
    loss, _ = ops.rnnt_fwd(joint, targets, joint_lengths, target_lengths, bool(from_logits))
    return loss


def transducer_forward_score_fg(f, g, targets, joint_lengths, target_lengths):
    """RNN-T negative log-likelihood per utterance of the additive joint the reference builds at
    ha/recognizer.py:114, `joint = f[:, :, None, :] + g[:, None, :, :]`, WITHOUT building it.

    f (N,T,K) = classifier(features), g (N,U+1,K) = lm outputs: raw float32 CUDA logits (the log-softmax over
    K is fused).  Equals transducer_forward_score(joint.log_softmax(-1), ...) and back-propagates
    d loss/d f = (d loss/d joint).sum(2), d loss/d g = (d loss/d joint).sum(1); the (N,T,U+1,K) tensor and
    its gradient (6.6 GB each at BASELINE config 4) never exist.
    """
    if functional.transforms_active():          # torch.func.grad / vmap: see functional.py
        return functional.rnnt_fg(f, g, targets, joint_lengths, target_lengths)
    loss, _ = ops.rnnt_fg_fwd(f, g, targets, joint_lengths, target_lengths)
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

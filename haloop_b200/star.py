"""Drop-in for ha/star.py: star / wildcard CTC (STC, Pratap et al. 2022).

    star_ctc_forward_score(emissions, targets, emission_lengths, target_lengths,
                           star_penalty=-0.5, animate=False) -> (N,)                 ha/star.py:65-163
"""
from . import functional, ops


def star_ctc_forward_score(emissions, targets, emission_lengths, target_lengths,
                           star_penalty=-0.5, animate=False, from_logits=False):
    """Star-CTC negative log-likelihood per utterance; see ctc_forward_score3 for the conventions.

    The (T,N,2V) star emission tensor of ha/star.py:8-49 is never materialised.  `animate` (a
    debugging printout of the reference, ha/star.py:150-152) is accepted and ignored.
    """
    if functional.transforms_active():          # torch.func.grad / vmap: see functional.py
        return functional.star(emissions, targets, emission_lengths, target_lengths, star_penalty, from_logits)
    loss, _ = ops.star_fwd(emissions, targets, emission_lengths, target_lengths,
                           float(star_penalty), bool(from_logits))
    return loss

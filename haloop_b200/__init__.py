"""haloop_b200 — B200 (sm_100a) alignment losses for haloop: CTC, star-CTC and RNN-T loss + gradient
behind the reference's own Python signatures (ha/ctc.py, ha/star.py, ha/transducer.py).

The hot path is hand-written CUDA in libha_b200.so (C ABI: include/ha_b200.h), bound with ctypes
and registered as PyTorch custom ops with autograd.  There is no CPU fallback.
"""
from .ctc import ctc_forward_score3, ctc_reduce_mean, ctc_loss
from .star import star_ctc_forward_score
from .transducer import transducer_forward_score, transducer_forward_score_fg, rnnt_loss
from .align import greedy_decode, greedy_decode_nested, ctc_viterbi_align, ctc_beam_search_decode_logits
from .head import linear_ctc_forward_score, linear_ctc_loss
from .recognizer import patch_haloop
from . import loop, sharding

__all__ = [
    "ctc_forward_score3", "ctc_reduce_mean", "ctc_loss", "star_ctc_forward_score",
    "transducer_forward_score", "transducer_forward_score_fg", "rnnt_loss", "greedy_decode", "greedy_decode_nested",
    "ctc_viterbi_align", "ctc_beam_search_decode_logits", "linear_ctc_forward_score", "linear_ctc_loss", "patch_haloop",
]
__version__ = "0.1.0"

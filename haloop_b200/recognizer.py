"""Adapter for the reference's call sites (ha/recognizer.py:61-82, 95-127).

The reference tree is never edited.  patch_haloop() rebinds the names ha.recognizer imported from
ha.ctc / ha.star / ha.transducer (ha/recognizer.py:6-8) and installs forward() methods that call
them, fixing the two call sites that are broken or disabled at HEAD:
  - ha/recognizer.py:71-72  F.ctc_loss live, ctc_forward_score3 commented out
  - ha/recognizer.py:79-80  star branch reads a non-existent self.star_penalty
  - ha/recognizer.py:116-126  transducer_forward_score under `if False:`
"""
import torch

from .ctc import ctc_forward_score3, ctc_reduce_mean
from .head import linear_ctc_forward_score
from .star import star_ctc_forward_score
from .transducer import transducer_forward_score, transducer_forward_score_fg, rnnt_loss


def temporal_classifier_forward(self, features, targets, input_lengths=None, target_lengths=None,
                                star_penalty=None, measure_entropy=False, drop_labels=False):
    """TemporalClassifier.forward (ha/recognizer.py:61-82) on the fused kernels.  CTC branch: the classifier itself is
    part of the op (haloop_b200.linear_ctc_forward_score: Linear -> log_softmax -> CTC on the tensor cores, the (N,T,C)
    logits never written) whenever the feature dimension is a multiple of 4 (`self.fused_head = False` turns it off);
    otherwise, and on the star branch, the classifier's raw logits go straight into the loss (log-softmax fused,
    from_logits=True)."""
    N, T = features.shape[0], features.shape[1]
    dev = features.device
    if input_lengths is None:
        input_lengths = torch.full((N,), T, dtype=torch.long, device=dev)
    if target_lengths is None:
        target_lengths = torch.full((N,), targets.shape[-1], dtype=torch.long, device=dev)
    with torch.autocast(device_type="cuda", enabled=False):
        feats = self.dropout(features).float()
        lin = self.classifier
        if (star_penalty is None and getattr(self, "fused_head", True) and type(lin) is torch.nn.Linear
                and feats.dim() == 3 and feats.shape[-1] % 4 == 0 and feats.is_cuda and targets.dim() == 2):
            losses = linear_ctc_forward_score(feats, lin.weight.float(), None if lin.bias is None else lin.bias.float(),
                                              targets, input_lengths, target_lengths)
            return (losses / target_lengths.to(losses.device).clamp_min(1)).mean(), {}
        logits = lin(feats).float()                                            # (N,T,C)
        logits = logits.permute(1, 0, 2)                                       # (T,N,C) view, no copy
        if star_penalty is None:
            losses = ctc_forward_score3(logits, targets, input_lengths, target_lengths, from_logits=True)
        else:
            losses = star_ctc_forward_score(logits, targets, input_lengths, target_lengths,
                                            star_penalty=star_penalty, from_logits=True)
        # the live call site is F.ctc_loss(reduction='mean') (ha/recognizer.py:71), which divides by
        # target_lengths.clamp_min(1): an empty transcript must not turn the batch loss into inf
        return (losses / target_lengths.to(losses.device).clamp_min(1)).mean(), {}


def transducer_forward(self, features, targets, input_lengths=None, target_lengths=None,
                       star_penalty=None, measure_entropy=False, drop_labels=False):
    """Transducer.forward (ha/recognizer.py:95-127) with the broadcast joint (:114) and rnnt_loss (:121-126)
    replaced by the joint-free lattice loss: the (N,T,U+1,C) tensor is never built."""
    N = features.shape[0]
    hidden = self.lm.init_hidden(N)
    lm_targets = torch.cat([targets.new_zeros((N, 1)), targets], dim=1)
    lm_outputs, _ = self.lm.forward_batch_first(lm_targets, hidden)
    features = self.classifier(self.dropout(features))
    with torch.autocast(device_type="cuda", enabled=False):
        losses = transducer_forward_score_fg(features.float(), lm_outputs.float(), targets,
                                             input_lengths, target_lengths)
    return losses.mean(), {}          # torchaudio's reduction='mean' is a plain batch mean


def patch_haloop(recognizer_module=None, patch_forward=True):
    """Rebind the hot-path names inside ha.recognizer (or the module given) to the CUDA versions.

    Returns the dict of the original attributes so the patch can be undone with unpatch_haloop().
    """
    if recognizer_module is None:
        import ha.recognizer as recognizer_module
    saved = {}

    def swap(obj, name, new):
        saved[(obj, name)] = getattr(obj, name, None)
        setattr(obj, name, new)

    swap(recognizer_module, "ctc_forward_score3", ctc_forward_score3)
    swap(recognizer_module, "ctc_reduce_mean", ctc_reduce_mean)
    swap(recognizer_module, "star_ctc_forward_score", star_ctc_forward_score)
    swap(recognizer_module, "transducer_forward_score", transducer_forward_score)
    if patch_forward:
        if hasattr(recognizer_module, "TemporalClassifier"):
            swap(recognizer_module.TemporalClassifier, "forward", temporal_classifier_forward)
        if hasattr(recognizer_module, "Transducer"):
            swap(recognizer_module.Transducer, "forward", transducer_forward)
    return saved


def unpatch_haloop(saved):
    for (obj, name), old in saved.items():
        if old is None:
            delattr(obj, name)
        else:
            setattr(obj, name, old)

"""Golden vectors for the call-site adapter (haloop_b200/recognizer.py): run the REAL reference modules
(ha.recognizer.TemporalClassifier from /root/reference, unmodified, float64, CPU) on seeded inputs and store
inputs, parameters, losses and parameter gradients in tests/golden/adapter_*.npz.

    python oracle/make_adapter_golden.py        # needs /root/reference (this container only)

TEST INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: tests/test_adapter.py rebuilds a module with the
same parameters from the fixture, patches it with patch_haloop() and compares.
Cases: the live CTC call site (ha/recognizer.py:71, F.ctc_loss 'mean') with an EMPTY transcript in the batch; the
star branch as the call site means it (ha/recognizer.py:78-81 with the star_penalty argument instead of the
non-existent self.star_penalty); the CTC term of CTCAttentionDecoder (ha/transformer.py:49-54: prompt token
stripped, weight 0.3).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import ha.recognizer as R                     # noqa: E402
from ha.ctc import ctc_reduce_mean            # noqa: E402
from ha.star import star_ctc_forward_score    # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def head_golden():
    """The fused head's own fixture: a wider TemporalClassifier (several 128-class / 32-feature tiles, ragged
    lengths, repeated labels), loss per utterance through the module's log_probs (ha/recognizer.py:43-46) and the
    reference's ctc_forward_score3, gradients w.r.t. features, weight and bias for a non-uniform grad_output."""
    from ha.ctc import ctc_forward_score3
    torch.manual_seed(99)
    N, T, D, V, U = 4, 150, 72, 300, 24
    g = torch.Generator().manual_seed(23)
    feats = torch.randn(N, T, D, generator=g, dtype=torch.float64).requires_grad_(True)
    tg = torch.randint(1, V, (N, U), generator=g)
    tg[1, 3:9] = tg[1, 3]                                   # repeated labels
    il = torch.tensor([150, 120, 97, 150]); tl = torch.tensor([24, 20, 11, 1])
    go = torch.tensor([1.0, 0.5, 2.0, 0.25], dtype=torch.float64)
    tc = R.TemporalClassifier(D, V).double()
    tc.dropout.p = 0.0
    losses = ctc_forward_score3(tc.log_probs(feats).permute(1, 0, 2), tg, il, tl)
    (losses * go).sum().backward()
    out = {"feats": feats.detach().numpy(), "targets": tg.numpy(), "in_len": il.numpy(), "tgt_len": tl.numpy(),
           "weight": tc.classifier.weight.detach().numpy(), "bias": tc.classifier.bias.detach().numpy(),
           "grad_out": go.numpy(), "loss": losses.detach().numpy(), "dh": feats.grad.numpy(),
           "dW": tc.classifier.weight.grad.numpy(), "db": tc.classifier.bias.grad.numpy()}
    np.savez_compressed(os.path.join(OUT, "head_temporal_classifier.npz"), **out)
    print("wrote head_temporal_classifier.npz", {k: v.shape for k, v in out.items()})


def main():
    torch.manual_seed(4321)
    N, T, D, V, U = 5, 40, 12, 9, 6
    g = torch.Generator().manual_seed(17)
    feats = torch.randn(N, T, D, generator=g, dtype=torch.float64)
    tg = torch.randint(1, V, (N, U), generator=g)
    il = torch.tensor([40, 37, 25, 31, 40]); tl = torch.tensor([6, 5, 0, 3, 6])       # utterance 2: empty transcript
    tg = tg * (torch.arange(U)[None, :] < tl[:, None])
    tc = R.TemporalClassifier(D, V).double()
    tc.dropout.p = 0.0
    W, b = tc.classifier.weight.detach().clone(), tc.classifier.bias.detach().clone()
    out = {"feats": feats.numpy(), "targets": tg.numpy(), "in_len": il.numpy(), "tgt_len": tl.numpy(),
           "weight": W.numpy(), "bias": b.numpy()}

    # (1) the live call site, unmodified forward
    loss, _ = tc(feats, tg, il, tl)
    loss.backward()
    out["ctc_loss"] = loss.item(); out["ctc_gw"] = tc.classifier.weight.grad.numpy().copy()
    out["ctc_gb"] = tc.classifier.bias.grad.numpy().copy()
    tc.zero_grad()

    # (2) the star branch as written, with the argument in place of self.star_penalty; lengths >= 1 (the
    #     reference divides by the raw target length)
    tl2 = tl.clamp_min(1)
    lp = tc.log_probs(feats).permute(1, 0, 2)
    losses = star_ctc_forward_score(lp, tg, il, tl2, star_penalty=-0.7)
    loss = ctc_reduce_mean(losses, tl2)
    loss.backward()
    out["star_penalty"] = -0.7; out["star_tgt_len"] = tl2.numpy()
    out["star_loss"] = loss.item(); out["star_gw"] = tc.classifier.weight.grad.numpy().copy()
    out["star_gb"] = tc.classifier.bias.grad.numpy().copy()
    tc.zero_grad()

    # (3) CTCAttentionDecoder's CTC term: condtargets carry one prompt token in front (ha/transformer.py:49-54)
    cond = torch.cat([torch.full((N, 1), 7), tg], dim=1); cl = tl + 1
    loss, _ = tc(feats, cond[:, 1:], il, cl - 1, None)
    loss = 0.3 * loss
    loss.backward()
    out["att_condtargets"] = cond.numpy(); out["att_cond_len"] = cl.numpy()
    out["att_loss"] = loss.item(); out["att_gw"] = tc.classifier.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "adapter_temporal_classifier.npz"), **out)
    head_golden()
    print("wrote adapter_temporal_classifier.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""patch_haloop() on a stand-in for ha.recognizer (the real module imports kaldialign / g2p_en / wandb,
absent here and on the GPU box; SURVEY 8f rank 3 asks for a stub).  The stand-in mirrors the attribute
structure of TemporalClassifier (ha/recognizer.py:35-46) and Transducer (ha/recognizer.py:84-89)."""
import types

import pytest
import torch
from torch import nn

import haloop_b200 as hb
from haloop_b200 import recognizer as adapter


class _LM(nn.Module):                      # ha/rnnlm.py Decoder, reduced to what Transducer.forward calls
    def __init__(self, vocab, dim=16):
        super().__init__()
        self.emb = nn.Embedding(vocab, dim)
        self.rnn = nn.GRU(dim, dim, batch_first=True)
        self.out = nn.Linear(dim, vocab)

    def init_hidden(self, n):
        return None

    def forward_batch_first(self, tokens, hidden):
        y, h = self.rnn(self.emb(tokens), hidden)
        return self.out(y), h


def _stub_module():
    m = types.ModuleType("ha_recognizer_stub")

    class TemporalClassifier(nn.Module):
        def __init__(self, feat_dim=12, vocab_size=9):
            super().__init__()
            self.classifier = nn.Linear(feat_dim, vocab_size)
            self.dropout = nn.Dropout(0.0)

        def forward(self, *a, **k):
            raise RuntimeError("reference forward")

    class Transducer(nn.Module):
        def __init__(self, feat_dim=12, vocab_size=9):
            super().__init__()
            self.classifier = nn.Linear(feat_dim, vocab_size)
            self.lm = _LM(vocab_size)
            self.dropout = nn.Dropout(0.0)

        def forward(self, *a, **k):
            raise RuntimeError("reference forward")

    m.TemporalClassifier, m.Transducer = TemporalClassifier, Transducer
    for name in ("ctc_forward_score3", "ctc_reduce_mean", "star_ctc_forward_score", "transducer_forward_score"):
        setattr(m, name, object())
    return m


def test_patch_and_unpatch_rebind_the_reference_names():
    m = _stub_module()
    before = {k: getattr(m, k) for k in ("ctc_forward_score3", "ctc_reduce_mean", "star_ctc_forward_score",
                                         "transducer_forward_score")}
    fwd = (m.TemporalClassifier.forward, m.Transducer.forward)
    saved = hb.patch_haloop(m)
    assert m.ctc_forward_score3 is hb.ctc_forward_score3 and m.ctc_reduce_mean is hb.ctc_reduce_mean
    assert m.star_ctc_forward_score is hb.star_ctc_forward_score
    assert m.transducer_forward_score is hb.transducer_forward_score
    assert m.TemporalClassifier.forward is adapter.temporal_classifier_forward
    assert m.Transducer.forward is adapter.transducer_forward
    adapter.unpatch_haloop(saved)
    assert all(getattr(m, k) is v for k, v in before.items())
    assert (m.TemporalClassifier.forward, m.Transducer.forward) == fwd


@pytest.mark.gpu
def test_patched_forwards_match_the_live_call_sites():
    """TemporalClassifier.forward vs F.ctc_loss(reduction='mean') (ha/recognizer.py:71) and Transducer.forward vs
    the broadcast joint + torchaudio rnnt_loss (ha/recognizer.py:114-126): same loss, same parameter gradients."""
    torchaudio = pytest.importorskip("torchaudio")
    import torch.nn.functional as F
    dev = torch.device("cuda")
    torch.manual_seed(1234)                    # module parameters come from the global generator
    m = _stub_module()
    saved = hb.patch_haloop(m)

    def close(a, b):                           # the library side of the comparison is plain fp32
        return (a - b).abs().max() <= 1e-4 + 1e-3 * b.abs().max()
    try:
        g = torch.Generator().manual_seed(3)
        N, T, D, V, U = 4, 30, 12, 9, 6
        feats = torch.randn(N, T, D, generator=g).to(dev)
        tg = torch.randint(1, V, (N, U), generator=g).to(dev)
        il = torch.tensor([30, 28, 20, 25]).to(dev); tl = torch.tensor([6, 5, 3, 6]).to(dev)

        tc = m.TemporalClassifier(D, V).to(dev)
        loss, _ = tc(feats, tg, il, tl)
        loss.backward()
        got = tc.classifier.weight.grad.clone(); tc.zero_grad()
        lp = tc.classifier(feats).log_softmax(-1).permute(1, 0, 2)
        ref = F.ctc_loss(lp.double(), tg, il, tl)               # reduction='mean': / target_lengths, batch mean
        ref.backward()
        assert abs(float(loss) / float(ref) - 1) < 1e-4
        assert close(got, tc.classifier.weight.grad)

        tr = m.Transducer(D, V).to(dev)
        loss, _ = tr(feats, tg, il, tl)
        loss.backward()
        got = [p.grad.clone() for p in tr.parameters()]; tr.zero_grad()
        lm_out, _ = tr.lm.forward_batch_first(torch.cat([tg.new_zeros((N, 1)), tg], 1), None)
        joint = tr.classifier(feats)[:, :, None, :] + lm_out[:, None, :, :]
        ref = torchaudio.functional.rnnt_loss(joint, tg.int(), il.int(), tl.int(), blank=0, reduction="mean",
                                              fused_log_softmax=True)
        ref.backward()
        assert abs(float(loss) / float(ref) - 1) < 1e-4
        for a, p in zip(got, tr.parameters()):
            assert close(a, p.grad)
    finally:
        adapter.unpatch_haloop(saved)

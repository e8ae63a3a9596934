"""CTC prefix beam search on the GPU (ha/beam.py:71-137) against the reference's own outputs (tests/golden/beam_*.npz,
graves_*.npz, written by oracle/make_beam_golden.py from the unmodified reference) and, at sizes the reference
cannot pin (its scores tie after a dozen frames), against the numpy restatement oracle/beam_oracle.py.
Hypotheses must match symbol for symbol; scores to 1e-5 relative + 1e-6 absolute (float32 logaddexp chains)."""
import numpy as np
import pytest
import torch

from conftest import golden_path

pytestmark = pytest.mark.gpu

REFERENCE_MODE = ["beam_small", "beam_short", "beam_peaky", "beam_blanky"]
GRAVES_MODE = ["graves_small", "graves_medium", "graves_peaky", "graves_blanky", "graves_wide"]


@pytest.fixture(scope="module")
def hb():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import haloop_b200
    return haloop_b200


@pytest.mark.parametrize("name", REFERENCE_MODE + GRAVES_MODE)
def test_beam_search_vs_reference_golden(hb, name):
    d = np.load(golden_path(name))
    graves = name.startswith("graves")
    for i in range(int(d["n"])):
        lp = torch.from_numpy(d[f"lp_{i}"]).cuda()
        for b in d["beams"]:
            seqs, sc = hb.ctc_beam_search_decode_logits(lp, beam_size=int(b), graves=graves)
            hl = d[f"len_{i}_{b}"]
            want = [d[f"hyp_{i}_{b}"][j, :hl[j]].tolist() for j in range(len(hl))]
            assert seqs == want, f"{name}[{i}] beam {b}"
            np.testing.assert_allclose(sc.cpu().numpy(), d[f"score_{i}_{b}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("T,K,beam,graves", [(120, 40, 8, True), (300, 64, 16, True), (64, 1024, 5, True),
                                             (9, 33, 16, False), (40, 7, 4, True)])
def test_beam_search_batch_vs_oracle(hb, T, K, beam, graves):
    """A ragged batch (input_lengths honoured, permuted view) against the numpy restatement, utterance by utterance."""
    from oracle import beam_oracle
    from haloop_b200 import ops
    g = torch.Generator().manual_seed(T * 131 + K)
    N = 5
    x = torch.randn(T, N, K, generator=g)
    x[:, :, 0] += 2.5 * (torch.rand(T, N, generator=g) < 0.6)
    lp = x.log_softmax(-1)
    il = torch.randint(max(T // 2, 1), T + 1, (N,), generator=g); il[0] = T
    hyp, hl, sc = ops.ctc_beam_search(lp.cuda().permute(1, 0, 2), il.cuda(), beam, not graves)
    hyp, hl, sc = hyp.cpu().numpy(), hl.cpu().numpy(), sc.cpu().numpy()
    for n in range(N):
        seqs, v = beam_oracle.ctc_beam_search(lp[:il[n], n].numpy(), beam, -np.inf if graves else 0.0, np.float32)
        for j, s in enumerate(seqs):
            if not np.isfinite(v[j]):
                continue
            assert hl[n, j] == len(s) and hyp[n, j, :len(s)].tolist() == s, f"utterance {n} beam {j}"
            assert (hyp[n, j, len(s):] == -1).all()
        np.testing.assert_allclose(sc[n, :len(v)], v, rtol=1e-5, atol=1e-5)


def test_beam_one_is_not_greedy_but_matches_the_oracle(hb):
    """beam_size=1 keeps the single best prefix per frame (not the argmax collapse)."""
    from oracle import beam_oracle
    g = torch.Generator().manual_seed(5)
    lp = torch.randn(50, 12, generator=g).log_softmax(-1)
    seqs, sc = hb.ctc_beam_search_decode_logits(lp.cuda(), beam_size=1, graves=True)
    want, v = beam_oracle.ctc_beam_search(lp.numpy(), 1, -np.inf, np.float32)
    assert seqs == want
    np.testing.assert_allclose(sc.cpu().numpy(), v, rtol=1e-5)


def test_beam_size_limits(hb):
    lp = torch.zeros(4, 8).log_softmax(-1).cuda()
    with pytest.raises(ValueError):
        hb.ctc_beam_search_decode_logits(lp, beam_size=17)
    with pytest.raises(ValueError):
        hb.ctc_beam_search_decode_logits(lp, beam_size=0)

"""Pins the CPU oracle (oracle/ha_oracle.c) to the reference.

tests/golden/*.npz were produced by oracle/make_golden.py running the unmodified
reference (ha/ctc.py:110-174, ha/star.py:65-163, ha/transducer.py:175-205) in
float64.  The oracle must reproduce loss and gradients to float64 round-off.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_path

LOSS_RTOL = 1e-11
GRAD_ATOL = 1e-10


def _cases(kind):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if name.startswith(kind + "_"):
            out.append(name)
    return out


def _regen_x(d, shape):
    import torch
    g = torch.Generator().manual_seed(int(d["seed"]))
    x = torch.randn(*shape, generator=g, dtype=torch.float32) * float(d.get("x_scale", 1.0))
    assert abs(float(x.double().sum()) - float(d["x_checksum"])) < 1e-6, "torch RNG drifted"
    return x.numpy()


def test_kat_appendix_d(oracle):
    """SURVEY.md Appendix D seeded known answers."""
    d = np.load(golden_path("kat_appendix_d"))
    loss, grad = oracle.ctc(d["x"], d["targets"], d["in_len"], d["tgt_len"])
    np.testing.assert_allclose(loss, [12.07453682, 11.34123406, 8.05409432], atol=1e-8)
    np.testing.assert_allclose(loss, d["ctc_loss"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(grad, d["ctc_grad"], atol=GRAD_ATOL)
    np.testing.assert_allclose(grad[0, 0], [-0.34710238, 0.17580602, 0.18838741, 0.34119272,
                                            -0.42690064, 0.06861688], atol=1e-8)
    assert abs(np.abs(grad).sum() - 28.920193629772072) < 1e-9
    assert abs(oracle.ctc_reduce_mean(loss, d["tgt_len"]) - 3.6086975725206556) < 1e-12
    loss, grad = oracle.star(d["x"], d["targets"], d["in_len"], d["tgt_len"], star_penalty=-0.5)
    np.testing.assert_allclose(loss, [8.56281179, 6.97394701, 6.39507674], atol=1e-8)
    np.testing.assert_allclose(grad, d["star_grad"], atol=GRAD_ATOL)
    assert abs(np.abs(grad).sum() - 17.561906477518058) < 1e-9
    loss, grad = oracle.rnnt(d["joint"], d["rnnt_targets"], d["rnnt_in_len"], d["rnnt_tgt_len"])
    np.testing.assert_allclose(loss, [15.62174349, 10.61215748, 10.60625096], atol=1e-8)
    np.testing.assert_allclose(grad, d["rnnt_grad"], atol=GRAD_ATOL)
    assert abs(np.abs(grad).sum() - 37.148514074504035) < 1e-9


def _check(d, loss, grad, lpgrad_fn, batch_axis):
    np.testing.assert_allclose(loss, d["loss"], rtol=LOSS_RTOL)
    assert abs(np.abs(grad).sum() - float(d["grad_abs_sum"])) < 1e-7 * max(1.0, float(d["grad_abs_sum"]))
    if "grad" in d:
        np.testing.assert_allclose(grad, d["grad"], atol=GRAD_ATOL)
        if "lpgrad" in d:
            np.testing.assert_allclose(lpgrad_fn(), d["lpgrad"], atol=GRAD_ATOL)
    else:
        sub = np.take(grad, d["grad_rows"], axis=batch_axis)
        np.testing.assert_allclose(sub, d["grad_sub"], atol=2e-7)   # stored as float32


@pytest.mark.parametrize("name", _cases("ctc"))
def test_ctc_golden(oracle, name):
    d = np.load(golden_path(name))
    T, N, V, S = d["shape"]
    x = d["x"] if "x" in d else _regen_x(d, (T, N, V))
    args = (d["targets"], d["in_len"], d["tgt_len"])
    loss, grad = oracle.ctc(x, *args)
    lp = x.astype(np.float64) - np.log(np.exp(x.astype(np.float64)).sum(-1, keepdims=True))
    _check(d, loss, grad, lambda: oracle.ctc(lp, *args, from_logits=False)[1], 1)
    assert abs(oracle.ctc_reduce_mean(loss, d["tgt_len"]) - float(d["reduce_mean"])) < 1e-10
    # rows beyond the input length carry exactly zero gradient (SURVEY §8b [probe])
    for n in range(N):
        assert not grad[int(d["in_len"][n]):, n].any()


@pytest.mark.parametrize("name", _cases("star"))
def test_star_golden(oracle, name):
    d = np.load(golden_path(name))
    T, N, V, S = d["shape"]
    x = d["x"] if "x" in d else _regen_x(d, (T, N, V))
    args = (d["targets"], d["in_len"], d["tgt_len"])
    pen = float(d["star_penalty"])
    loss, grad = oracle.star(x, *args, star_penalty=pen)
    lp = x.astype(np.float64) - np.log(np.exp(x.astype(np.float64)).sum(-1, keepdims=True))
    _check(d, loss, grad, lambda: oracle.star(lp, *args, star_penalty=pen, from_logits=False)[1], 1)


@pytest.mark.parametrize("name", _cases("rnnt"))
def test_rnnt_golden(oracle, name):
    d = np.load(golden_path(name))
    N, T, U, V = d["shape"]
    x = d["x"] if "x" in d else _regen_x(d, (N, T, U + 1, V))
    args = (d["targets"], d["in_len"], d["tgt_len"])
    loss, grad = oracle.rnnt(x, *args)
    lp = x.astype(np.float64) - np.log(np.exp(x.astype(np.float64)).sum(-1, keepdims=True))
    _check(d, loss, grad, lambda: oracle.rnnt(lp, *args, from_logits=False)[1], 0)


@pytest.mark.parametrize("name", _cases("rnntfg"))
def test_rnnt_joint_free_golden(oracle, name):
    """the reference's additive joint (ha/recognizer.py:114) differentiated w.r.t. its factors f and g"""
    d = np.load(golden_path(name))
    loss, gf, gg = oracle.rnnt_fg(d["f"], d["g"], d["targets"], d["in_len"], d["tgt_len"])
    np.testing.assert_allclose(loss, d["loss"], rtol=1e-11)
    assert np.abs(gf - d["grad_f"]).max() < 1e-10
    assert np.abs(gg - d["grad_g"]).max() < 1e-10


def test_ctc_matches_torch_ctc_loss(oracle):
    """ha/ctc.py:205-238 prints agreement with F.ctc_loss; fp64 agreement is exact."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    T, N, V, S = 70, 5, 17, 9
    x = torch.randn(T, N, V, generator=g, dtype=torch.float64, requires_grad=True)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.tensor([70, 60, 50, 41, 33]); tl = torch.tensor([9, 8, 3, 1, 9])
    ref = F.ctc_loss(x.log_softmax(-1), tg, il, tl, reduction="none")
    ref.sum().backward()
    loss, grad = oracle.ctc(x.detach().numpy(), tg.numpy(), il.numpy(), tl.numpy())
    np.testing.assert_allclose(loss, ref.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(grad, x.grad.numpy(), atol=1e-11)


def test_rnnt_matches_torchaudio(oracle):
    """ha/transducer.py:209-234 (test_batched): same shapes/seed, vs torchaudio rnnt_loss."""
    import torch
    torchaudio = pytest.importorskip("torchaudio")
    torch.manual_seed(42)
    f = torch.randn(13, 7, 6); gq = torch.randn(13, 5, 6)
    tg = torch.randint(0, 6, (13, 4))
    il = torch.tensor([7] * 13, dtype=torch.int32); tl = torch.tensor([4] * 13, dtype=torch.int32)
    joint = (f[:, :, None, :] + gq[:, None, :, :]).log_softmax(-1)
    ref = torchaudio.functional.rnnt_loss(joint, tg.to(torch.int32), il, tl, blank=0,
                                          reduction="none", fused_log_softmax=False)
    loss, _ = oracle.rnnt(joint.numpy(), tg.numpy(), il.numpy(), tl.numpy(), from_logits=False,
                          want_grad=False)
    np.testing.assert_allclose(loss, ref.numpy(), rtol=2e-6)


def test_edge_cases(oracle):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((6, 3, 5))
    tg = np.array([[1, 2, 3], [1, 1, 1], [2, 0, 0]])
    # utt0 feasible, utt1 infeasible (needs 5 frames for A_A_A, has 4), utt2 empty target
    il = np.array([6, 4, 5]); tl = np.array([3, 3, 0])
    loss, grad = oracle.ctc(x, tg, il, tl)
    assert np.isfinite(loss[0]) and np.isinf(loss[1]) and loss[1] > 0
    assert not grad[:, 1].any()
    lp = x - np.log(np.exp(x).sum(-1, keepdims=True))
    assert abs(loss[2] + lp[:5, 2, 0].sum()) < 1e-12
    # rows sum to zero through the softmax, and to -1 at the log-prob boundary
    assert np.abs(grad[:, 0].sum(-1)).max() < 1e-12
    _, glp = oracle.ctc(lp, tg, il, tl, from_logits=False)
    np.testing.assert_allclose(glp[:6, 0].sum(-1), -1.0, atol=1e-12)
    # grad_out scaling
    _, g2 = oracle.ctc(x, tg, il, tl, grad_out=np.array([2.0, 1.0, -3.0]))
    np.testing.assert_allclose(g2[:, 0], 2.0 * grad[:, 0], atol=1e-14)
    np.testing.assert_allclose(g2[:, 2], -3.0 * grad[:, 2], atol=1e-14)
    # star: T=1, y=[1] gives exactly -lp[0,1] (SURVEY Appendix A.2 [probe])
    l, _ = oracle.star(lp[:1, :1], np.array([[1]]), np.array([1]), np.array([1]), from_logits=False)
    assert abs(l[0] + lp[0, 0, 1]) < 1e-12


def test_finite_difference(oracle):
    """Closed-form gradients vs central differences of the oracle's own loss."""
    rng = np.random.default_rng(1)
    T, N, V, S = 9, 2, 5, 3
    x = rng.standard_normal((T, N, V))
    tg = np.array([[1, 1, 2], [3, 4, 0]]); il = np.array([9, 7]); tl = np.array([3, 2])
    for fn in (oracle.ctc, lambda *a, **k: oracle.star(*a, star_penalty=-0.3, **k)):
        _, grad = fn(x, tg, il, tl)
        for idx in [(0, 0, 1), (3, 0, 0), (5, 1, 4), (8, 0, 2), (8, 1, 1)]:
            xp = x.copy(); xp[idx] += 1e-6
            xm = x.copy(); xm[idx] -= 1e-6
            fd = (fn(xp, tg, il, tl, want_grad=False)[0].sum() -
                  fn(xm, tg, il, tl, want_grad=False)[0].sum()) / 2e-6
            assert abs(fd - grad[idx]) < 1e-7
    j = rng.standard_normal((2, 5, 4, 6))
    tg = np.array([[1, 2, 2], [5, 0, 0]]); il = np.array([5, 4]); tl = np.array([3, 1])
    _, grad = oracle.rnnt(j, tg, il, tl)
    for idx in [(0, 0, 0, 0), (0, 2, 1, 2), (1, 3, 1, 0), (1, 0, 0, 5), (0, 4, 3, 0)]:
        jp = j.copy(); jp[idx] += 1e-6
        jm = j.copy(); jm[idx] -= 1e-6
        fd = (oracle.rnnt(jp, tg, il, tl, want_grad=False)[0].sum() -
              oracle.rnnt(jm, tg, il, tl, want_grad=False)[0].sum()) / 2e-6
        assert abs(fd - grad[idx]) < 1e-7
    # padded nodes carry zero gradient
    assert not grad[1, 4:].any() and not grad[1, :, 2:].any()


def test_greedy_matches_reference_semantics(oracle):
    """ha/recognizer.py:48-59 restated with torch: max -> unique_consecutive -> drop 0."""
    import torch
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 25, 7, generator=g).log_softmax(-1)
    x[0, 3] = x[0, 3, 2]   # a full tie: first index (0) must win, as torch.max does
    ali, sc, hyp, hl = oracle.greedy(x.numpy())
    scores, alignments = x.max(dim=-1)
    np.testing.assert_array_equal(ali, alignments.numpy())
    np.testing.assert_array_equal(sc, scores.double().numpy())
    for n in range(4):
        ref = [int(i) for i in torch.unique_consecutive(alignments[n]) if i]
        assert list(hyp[n, :hl[n]]) == ref


def test_viterbi_consistency(oracle):
    rng = np.random.default_rng(7)
    T, N, V, S = 30, 3, 6, 5
    x = rng.standard_normal((T, N, V)).astype(np.float32)
    lp = (x - np.log(np.exp(x.astype(np.float64)).sum(-1, keepdims=True))).astype(np.float32)
    tg = rng.integers(1, V, (N, S)); il = np.array([30, 22, 11]); tl = np.array([5, 3, 5])
    ali, sc = oracle.ctc_viterbi(lp, tg, il, tl)
    loss, _ = oracle.ctc(lp, tg, il, tl, from_logits=False, want_grad=False)
    for n in range(N):
        path = ali[n, :il[n]]
        assert (ali[n, il[n]:] == -1).all()
        # collapsing the best path gives back the target
        col = [int(c) for i, c in enumerate(path) if c != 0 and (i == 0 or c != path[i - 1])]
        # repeated labels need a blank between them, so collapse-by-blank:
        col2, prev = [], 0
        for c in path:
            if c != 0 and c != prev:
                col2.append(int(c))
            prev = c
        assert col2 == list(tg[n, :tl[n]])
        s = sum(float(lp[t, n, path[t]]) for t in range(il[n]))
        assert abs(s - sc[n]) < 1e-3
        assert sc[n] <= -loss[n] + 1e-4     # best path <= total


# ------------------------------------------------------------------ prefix beam search (ha/beam.py:71-137) ---
BEAM_GOLDENS = ["beam_small", "beam_short", "beam_peaky", "beam_blanky",
                "graves_small", "graves_medium", "graves_peaky", "graves_blanky", "graves_wide"]


@pytest.mark.parametrize("name", BEAM_GOLDENS)
def test_beam_oracle_vs_reference_golden(name):
    """The numpy restatement reproduces the unmodified reference hypothesis for hypothesis (both modes)."""
    from oracle import beam_oracle
    from conftest import golden_path
    d = np.load(golden_path(name))
    assert int(d["n"]) >= 1
    ext_blank = -np.inf if name.startswith("graves") else 0.0
    for i in range(int(d["n"])):
        for b in d["beams"]:
            seqs, sc = beam_oracle.ctc_beam_search(d[f"lp_{i}"], int(b), ext_blank)
            hl = d[f"len_{i}_{b}"]
            assert [len(s) for s in seqs] == hl.tolist()
            for j, s in enumerate(seqs):
                assert s == d[f"hyp_{i}_{b}"][j, :hl[j]].tolist()
            np.testing.assert_allclose(sc, d[f"score_{i}_{b}"], rtol=1e-5 if name.startswith("graves") else 1e-10,
                                       atol=1e-6 if name.startswith("graves") else 1e-12)


# ------------------------------------------------------------ classifier head + CTC (ha/recognizer.py:43-46) ---
def test_head_oracle_vs_reference_module_golden(oracle):
    """oracle.head_ctc against the unmodified TemporalClassifier.log_probs + ctc_forward_score3 in float64."""
    from conftest import golden_path
    d = np.load(golden_path("head_temporal_classifier"))
    loss, dh, dW, db = oracle.head_ctc(d["feats"], d["weight"], d["bias"], d["targets"], d["in_len"], d["tgt_len"], d["grad_out"])
    np.testing.assert_allclose(loss, d["loss"], rtol=1e-11)
    np.testing.assert_allclose(dh, d["dh"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(dW, d["dW"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(db, d["db"], rtol=1e-9, atol=1e-11)

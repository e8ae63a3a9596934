"""Parity of the CUDA hot path (through the Python API -> custom op -> C ABI) against the golden
vectors of the reference (float64 runs of ha/ctc.py, ha/star.py, ha/transducer.py) and against the
CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): loss within 1e-4 relative, logit gradients within 1e-5
absolute, both against float64 gold; the reference's own float32 deviation (golden *_ref32_*) is
2e-6 .. 8e-5 on the same cases, i.e. the bar is tighter than the reference's fp32 noise floor.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_path

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_ATOL = 1e-5


@pytest.fixture(scope="module")
def hb():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import haloop_b200
    return haloop_b200


def dev():
    return torch.device("cuda:0")


def _cases(kind):
    return [os.path.basename(p)[:-4] for p in sorted(glob.glob(os.path.join(GOLDEN, kind + "_*.npz")))]


def _x_of(d, shape):
    if "x" in d:
        return torch.from_numpy(d["x"])
    g = torch.Generator().manual_seed(int(d["seed"]))
    x = torch.randn(*shape, generator=g, dtype=torch.float32) * float(d["x_scale"] if "x_scale" in d else 1.0)
    assert abs(float(x.double().sum()) - float(d["x_checksum"])) < 1e-6, "torch RNG drifted"
    return x


def _assert_loss(loss, ref):
    loss = loss.detach().double().cpu().numpy()
    np.testing.assert_allclose(loss, ref, rtol=LOSS_RTOL)


def _grad_check(grad, d, batch_axis):
    g = grad.double().cpu().numpy()
    if "grad" in d:
        err = np.abs(g - d["grad"]).max()
    else:
        err = np.abs(np.take(g, d["grad_rows"], axis=batch_axis) - d["grad_sub"]).max()
        assert abs(np.abs(g).sum() / float(d["grad_abs_sum"]) - 1) < 1e-5
    assert err < GRAD_ATOL, f"grad abs err {err:.3e} (reference fp32 itself: {float(d['ref32_grad_dev']):.3e})"


@pytest.mark.parametrize("name", _cases("ctc"))
def test_ctc_golden(hb, name):
    d = np.load(golden_path(name))
    T, N, V, S = d["shape"]
    x = _x_of(d, (T, N, V)).to(dev()).requires_grad_(True)
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("targets", "in_len", "tgt_len"))
    loss = hb.ctc_forward_score3(x, tg, il, tl, from_logits=True)
    _assert_loss(loss, d["loss"])
    assert abs(float(hb.ctc_reduce_mean(loss, tl)) / float(d["reduce_mean"]) - 1) < LOSS_RTOL
    loss.sum().backward()
    _grad_check(x.grad, d, 1)
    for n in range(N):
        assert not x.grad[int(il[n]):, n].any(), "rows beyond the input length must be exactly zero"
    if "lpgrad" in d:
        # reference contract: log-probs in, autograd gives -occupancy at the emissions boundary
        lp = _x_of(d, (T, N, V)).double().log_softmax(-1).float().to(dev()).requires_grad_(True)
        l2 = hb.ctc_forward_score3(lp, tg, il, tl)
        _assert_loss(l2, d["loss"])
        l2.sum().backward()
        assert np.abs(lp.grad.double().cpu().numpy() - d["lpgrad"]).max() < GRAD_ATOL


@pytest.mark.parametrize("name", _cases("star"))
def test_star_golden(hb, name):
    d = np.load(golden_path(name))
    T, N, V, S = d["shape"]
    pen = float(d["star_penalty"])
    x = _x_of(d, (T, N, V)).to(dev()).requires_grad_(True)
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("targets", "in_len", "tgt_len"))
    loss = hb.star_ctc_forward_score(x, tg, il, tl, star_penalty=pen, from_logits=True)
    _assert_loss(loss, d["loss"])
    loss.sum().backward()
    _grad_check(x.grad, d, 1)
    if "lpgrad" in d:
        lp = _x_of(d, (T, N, V)).double().log_softmax(-1).float().to(dev()).requires_grad_(True)
        l2 = hb.star_ctc_forward_score(lp, tg, il, tl, star_penalty=pen)
        _assert_loss(l2, d["loss"])
        l2.sum().backward()
        assert np.abs(lp.grad.double().cpu().numpy() - d["lpgrad"]).max() < GRAD_ATOL


@pytest.mark.parametrize("name", _cases("rnnt"))
def test_rnnt_golden(hb, name):
    d = np.load(golden_path(name))
    N, T, U, V = d["shape"]
    x = _x_of(d, (N, T, U + 1, V)).to(dev()).requires_grad_(True)
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("targets", "in_len", "tgt_len"))
    loss = hb.transducer_forward_score(x, tg, il, tl, from_logits=True)
    _assert_loss(loss, d["loss"])
    loss.sum().backward()
    _grad_check(x.grad, d, 0)
    if "lpgrad" in d:
        lp = _x_of(d, (N, T, U + 1, V)).double().log_softmax(-1).float().to(dev()).requires_grad_(True)
        l2 = hb.transducer_forward_score(lp, tg.int(), il.int(), tl.int())     # int32, as ha/transducer.py:217-218
        _assert_loss(l2, d["loss"])
        l2.sum().backward()
        assert np.abs(lp.grad.double().cpu().numpy() - d["lpgrad"]).max() < GRAD_ATOL


def test_kat_appendix_d(hb):
    d = np.load(golden_path("kat_appendix_d"))
    x = torch.from_numpy(d["x"]).float().to(dev())
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("targets", "in_len", "tgt_len"))
    _assert_loss(hb.ctc_forward_score3(x, tg, il, tl, from_logits=True), d["ctc_loss"])
    _assert_loss(hb.star_ctc_forward_score(x, tg, il, tl, star_penalty=-0.5, from_logits=True), d["star_loss"])
    j = torch.from_numpy(d["joint"]).float().to(dev())
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("rnnt_targets", "rnnt_in_len", "rnnt_tgt_len"))
    _assert_loss(hb.transducer_forward_score(j, tg, il, tl, from_logits=True), d["rnnt_loss"])


def _rand_ctc(seed, T, N, V, S, var=True, scale=1.0, repeats=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g) * scale
    tg = torch.randint(1, 3 if repeats else V, (N, S), generator=g)
    if var:
        il = torch.randint(T // 2, T + 1, (N,), generator=g); il[0] = T
        tl = torch.randint(S // 2, S + 1, (N,), generator=g); tl[0] = S
    else:
        il = torch.full((N,), T); tl = torch.full((N,), S)
    return x, tg, il, tl


@pytest.mark.parametrize("cfg", [
    dict(T=300, N=5, V=64, S=100),                 # 4 slots
    dict(T=700, N=3, V=128, S=300),                # 10 slots -> J=12 (the C2 target length)
    dict(T=97, N=7, V=37, S=11),                   # V % 4 != 0: scalar (non-bulk) path
    dict(T=400, N=4, V=32, S=40, scale=5.0),       # peaky posteriors
    dict(T=260, N=4, V=8, S=120, repeats=True),    # many repeated labels, T barely feasible for some
])
def test_ctc_vs_oracle(hb, oracle, cfg):
    c = dict(cfg)
    x, tg, il, tl = _rand_ctc(100 + c["T"], c.pop("T"), c.pop("N"), c.pop("V"), c.pop("S"), **c)
    go = torch.linspace(0.5, 2.0, x.shape[1])
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), grad_out=go.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    (loss * go.to(dev())).sum().backward()
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL)
    assert (np.isinf(lo) == np.isinf(ol)).all()
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("cfg", [
    dict(T=300, N=5, V=64, S=60),
    dict(T=500, N=3, V=128, S=200),                # 7 slots -> J=8 (the C3 target length)
    dict(T=97, N=6, V=37, S=11),
    dict(T=300, N=4, V=32, S=40, scale=4.0),
    dict(T=200, N=4, V=8, S=50, repeats=True),
])
def test_star_vs_oracle(hb, oracle, cfg):
    c = dict(cfg)
    x, tg, il, tl = _rand_ctc(200 + c["T"], c.pop("T"), c.pop("N"), c.pop("V"), c.pop("S"), **c)
    for n in range(tg.shape[0]):
        tg[n, tl[n]:] = 0                          # Collator pads with 0 (ha/loop.py:40)
    go = torch.linspace(0.5, 2.0, x.shape[1])
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5, grad_out=go.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), star_penalty=-0.5,
                                     from_logits=True)
    (loss * go.to(dev())).sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("cfg", [
    dict(N=3, T=50, U=20, V=64),
    dict(N=2, T=33, U=70, V=32),                   # U+1 spans 3 warps
    dict(N=4, T=61, U=9, V=37),                    # V % 4 != 0
    dict(N=2, T=150, U=100, V=16),                 # the C4 lattice width
])
def test_rnnt_vs_oracle(hb, oracle, cfg):
    N, T, U, V = cfg["N"], cfg["T"], cfg["U"], cfg["V"]
    g = torch.Generator().manual_seed(300 + T)
    x = torch.randn(N, T, U + 1, V, generator=g)
    tg = torch.randint(0, V, (N, U), generator=g)           # label 0 allowed, as ha/transducer.py:216
    il = torch.randint(T // 2, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (N,), generator=g); tl[0] = U
    go = torch.linspace(0.5, 2.0, N)
    ol, og = oracle.rnnt(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), grad_out=go.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.transducer_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    (loss * go.to(dev())).sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("S", [1, 2, 33, 64, 65, 129, 200, 333, 420, 520, 650, 800, 1000])
def test_ctc_target_length_sweep(hb, oracle, S):
    """every (slots per warp, warps per side) instantiation of the trellis kernel the launcher can pick"""
    T = S + S // 3 + 24
    x, tg, il, tl = _rand_ctc(900 + S, T, 2, 16, S)
    tl[0] = S; il[0] = T
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    assert (np.isinf(lo) == np.isinf(ol)).all()
    np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL)
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("S", [1, 33, 65, 100, 129, 260, 400, 511])
def test_star_target_length_sweep(hb, oracle, S):
    T = S + S // 3 + 24
    x, tg, il, tl = _rand_ctc(950 + S, T, 2, 16, S)
    tl[0] = S; il[0] = T
    for n in range(tg.shape[0]):
        tg[n, tl[n]:] = 0
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-1.25)
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), star_penalty=-1.25,
                                     from_logits=True)
    loss.sum().backward()
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    assert (np.isinf(lo) == np.isinf(ol)).all()
    np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL)
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("cfg", [dict(N=2, T=30, U=200, V=8), dict(N=2, T=6, U=520, V=4)])
def test_rnnt_wide_lattice(hb, oracle, cfg):
    """U+1 > 128 (the 1024-thread lattice instantiation) and U+1 > 512 (alpha and beta share the threads)"""
    N, T, U, V = cfg["N"], cfg["T"], cfg["U"], cfg["V"]
    g = torch.Generator().manual_seed(400 + U)
    x = torch.randn(N, T, U + 1, V, generator=g)
    tg = torch.randint(0, V, (N, U), generator=g)
    il = torch.tensor([T, max(1, T - 2)]); tl = torch.tensor([U, U - 7])
    ol, og = oracle.rnnt(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.transducer_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    err = np.abs(xd.grad.double().cpu().numpy() - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"


@pytest.mark.parametrize("name", _cases("rnntfg"))
def test_rnnt_joint_free_golden(hb, name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    f = torch.from_numpy(d["f"]).to(dev()).requires_grad_(True)
    g = torch.from_numpy(d["g"]).to(dev()).requires_grad_(True)
    tg, il, tl = (torch.from_numpy(d[k]).to(dev()) for k in ("targets", "in_len", "tgt_len"))
    loss = hb.transducer_forward_score_fg(f, g, tg, il, tl)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), d["loss"], rtol=LOSS_RTOL)
    assert np.abs(f.grad.double().cpu().numpy() - d["grad_f"]).max() < GRAD_ATOL
    assert np.abs(g.grad.double().cpu().numpy() - d["grad_g"]).max() < GRAD_ATOL


@pytest.mark.parametrize("cfg", [
    dict(N=3, T=50, U=20, V=64),
    dict(N=2, T=130, U=70, V=33),                  # several tiles in every GEMM, V % 4 != 0
    dict(N=4, T=61, U=9, V=200, zero=True),        # label 0 in the targets
    dict(N=2, T=200, U=100, V=96, scale=3.0),      # the C4 lattice width, peaky
    dict(N=2, T=300, U=100, V=256),                # tensor-core path: 128-column gradient tiles, 3 row tiles
    dict(N=2, T=40, U=300, V=48, scale=2.0),       # tensor-core path: U+1 > 256 (two E tiles), 16-column gradient tiles
    dict(N=3, T=129, U=15, V=1024, zero=True),     # tensor-core path: K = V = 1024, label 0
])
def test_rnnt_joint_free_vs_oracle(hb, oracle, cfg):
    N, T, U, V = cfg["N"], cfg["T"], cfg["U"], cfg["V"]
    gen = torch.Generator().manual_seed(500 + T)
    f = torch.randn(N, T, V, generator=gen) * cfg.get("scale", 1.0)
    g = torch.randn(N, U + 1, V, generator=gen) * cfg.get("scale", 1.0)
    tg = torch.randint(0 if cfg.get("zero") else 1, V, (N, U), generator=gen)
    il = torch.randint(T // 2, T + 1, (N,), generator=gen); il[0] = T
    tl = torch.randint(U // 2, U + 1, (N,), generator=gen); tl[0] = U
    go = torch.linspace(0.5, 2.0, N)
    ol, ogf, ogg = oracle.rnnt_fg(f.numpy(), g.numpy(), tg.numpy(), il.numpy(), tl.numpy(), grad_out=go.numpy())
    fd = f.to(dev()).requires_grad_(True); gd = g.to(dev()).requires_grad_(True)
    loss = hb.transducer_forward_score_fg(fd, gd, tg.to(dev()), il.to(dev()), tl.to(dev()))
    (loss * go.to(dev())).sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    # a gradient w.r.t. f sums U+1 joint gradients and one w.r.t. g sums T of them (the blank column reaches
    # magnitudes of ~T/2, where an fp32 ulp is already 4e-6): 1e-5 absolute plus 2e-6 relative
    for got, ref in ((fd.grad, ogf), (gd.grad, ogg)):
        err = np.abs(got.double().cpu().numpy() - ref)
        assert (err <= GRAD_ATOL + 2e-6 * np.abs(ref)).all(), f"{err.max():.3e}"


def test_rnnt_joint_free_equals_joint_path(hb):
    """same loss and (reduced) gradients as the materialised-joint op on the same GPU"""
    gen = torch.Generator().manual_seed(77)
    N, T, U, V = 3, 40, 12, 48
    f = torch.randn(N, T, V, generator=gen).to(dev()).requires_grad_(True)
    g = torch.randn(N, U + 1, V, generator=gen).to(dev()).requires_grad_(True)
    tg = torch.randint(1, V, (N, U), generator=gen).to(dev())
    il = torch.tensor([T, T - 3, T // 2]).to(dev()); tl = torch.tensor([U, U - 1, U // 2]).to(dev())
    l1 = hb.transducer_forward_score_fg(f, g, tg, il, tl)
    l1.sum().backward()
    gf1, gg1 = f.grad.clone(), g.grad.clone()
    f.grad = None; g.grad = None
    joint = f[:, :, None, :] + g[:, None, :, :]
    l2 = hb.transducer_forward_score(joint, tg, il, tl, from_logits=True)
    l2.sum().backward()
    torch.testing.assert_close(l1, l2, rtol=1e-5, atol=1e-4)
    assert (gf1 - f.grad).abs().max() < GRAD_ATOL and (gg1 - g.grad).abs().max() < GRAD_ATOL


def test_permuted_view_and_strided_grad(hb, oracle):
    """ha/recognizer.py:70: the loss sees logits.permute(1,0,2) of an (N,T,C) buffer; no copy is made
    and the gradient comes back with the same strides."""
    x, tg, il, tl = _rand_ctc(7, 120, 6, 48, 20)
    base = x.permute(1, 0, 2).contiguous().to(dev()).requires_grad_(True)     # (N,T,C) leaf
    view = base.permute(1, 0, 2)
    assert not view.is_contiguous()
    loss = hb.ctc_forward_score3(view, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    assert np.abs(base.grad.permute(1, 0, 2).double().cpu().numpy() - og).max() < GRAD_ATOL


def test_edge_cases(hb, oracle):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 4, 8, generator=g)
    tg = torch.tensor([[1, 2, 3], [1, 1, 1], [2, 0, 0], [4, 5, 0]])
    il = torch.tensor([6, 4, 5, 1]); tl = torch.tensor([3, 3, 0, 1])   # ok, infeasible, empty target, T=1
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    lo = loss.detach().cpu().numpy()
    assert np.isinf(lo[1]) and lo[1] > 0 and np.isinf(ol[1])
    np.testing.assert_allclose(lo[[0, 2, 3]], ol[[0, 2, 3]], rtol=LOSS_RTOL)
    assert not xd.grad[:, 1].any(), "infeasible utterance: zero gradient"
    assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL
    # out-of-range label / length -> NaN loss, zero gradient, other utterances untouched
    bad = tg.clone(); bad[0, 1] = 99
    xd2 = x.to(dev()).requires_grad_(True)
    l2 = hb.ctc_forward_score3(xd2, bad.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    l2[2:].sum().backward()
    assert torch.isnan(l2[0]) and not xd2.grad[:, 0].any()
    np.testing.assert_allclose(l2[2:].detach().cpu().numpy(), ol[2:], rtol=LOSS_RTOL)


def test_rnnt_live_call_site(hb):
    """rnnt_loss wrapper (ha/recognizer.py:121-126) vs torchaudio on the same GPU."""
    torchaudio = pytest.importorskip("torchaudio")
    g = torch.Generator().manual_seed(9)
    N, T, U, V = 4, 30, 8, 24
    x = torch.randn(N, T, U + 1, V, generator=g).to(dev())
    tg = torch.randint(1, V, (N, U), generator=g).to(dev())
    il = torch.tensor([30, 21, 30, 17]).to(dev()); tl = torch.tensor([8, 8, 3, 5]).to(dev())
    a = x.clone().requires_grad_(True)
    la = hb.rnnt_loss(a, tg.int(), il.int(), tl.int(), blank=0, reduction="mean", fused_log_softmax=True)
    la.backward()
    b = x.clone().requires_grad_(True)
    lb = torchaudio.functional.rnnt_loss(b, tg.int(), il.int(), tl.int(), blank=0, reduction="mean",
                                         fused_log_softmax=True)
    lb.backward()
    assert abs(float(la) / float(lb) - 1) < 1e-5
    assert (a.grad - b.grad).abs().max() < 1e-5


def test_ctc_live_call_site(hb):
    """ctc_loss wrapper (ha/recognizer.py:71) vs F.ctc_loss(reduction='mean') in float64 on the GPU."""
    x, tg, il, tl = _rand_ctc(11, 150, 6, 40, 25)
    a = x.to(dev()).requires_grad_(True)
    la = hb.ctc_loss(a, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    la.backward()
    b = x.double().to(dev()).requires_grad_(True)
    lb = torch.nn.functional.ctc_loss(b.log_softmax(-1), tg.to(dev()), il.to(dev()), tl.to(dev()))
    lb.backward()
    assert abs(float(la) / float(lb) - 1) < 1e-5
    assert (a.grad.double() - b.grad).abs().max() < 1e-6      # scaled by 1/(N*L): tighter in absolute terms


def test_greedy_bit_exact(hb, oracle):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 300, 40, generator=g).log_softmax(-1)
    x[0, 3] = x[0, 3, 2]                # a full tie: index 0 must win
    x[1, 7, 5] = x[1, 7].max()          # a two-way tie
    il = torch.tensor([300, 250, 1, 0, 299])
    hyp, hl, ali, sc = hb.greedy_decode(x.to(dev()))
    scores, alignments = x.max(dim=-1)
    assert torch.equal(ali.cpu(), alignments) and torch.equal(sc.cpu(), scores)
    for n in range(5):
        ref = [int(i) for i in torch.unique_consecutive(alignments[n]) if i]
        assert hyp[n, :hl[n]].tolist() == ref
        assert (hyp[n, hl[n]:] == -1).all()
    o_ali, o_sc, o_hyp, o_hl = oracle.greedy(x.numpy(), il.numpy())
    hyp, hl, ali, sc = hb.greedy_decode(x.to(dev()), il.to(dev()))
    assert np.array_equal(hyp.cpu().numpy(), o_hyp) and np.array_equal(hl.cpu().numpy(), o_hl)
    assert np.array_equal(ali.cpu().numpy(), o_ali)


def test_viterbi_bit_exact(hb, oracle):
    x, tg, il, tl = _rand_ctc(21, 200, 6, 30, 40)
    lp = x.log_softmax(-1)
    ali, sc = hb.ctc_viterbi_align(lp.to(dev()), tg.to(dev()), il.to(dev()), tl.to(dev()))
    o_ali, o_sc = oracle.ctc_viterbi(lp.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    assert np.array_equal(ali.cpu().numpy(), o_ali)
    assert np.array_equal(sc.cpu().numpy().view(np.int32), o_sc.view(np.int32)), "scores must match bit for bit"


# ------------------------------------------------------------------------------------ fuzz ---
def _fuzz_shapes(seed, n):
    import random
    r = random.Random(seed)
    out = []
    for _ in range(n):
        S = r.choice([0, 1, 2, 3, 7, 31, 32, 33, 40, 63, 64, 65, 90])
        T = max(1, S + r.choice([-2, 0, 1, 2, 5, 17, 40]))
        out.append(dict(T=T, N=r.choice([1, 2, 3, 5]), V=r.choice([2, 3, 5, 8, 17, 32, 100]), S=S,
                        scale=r.choice([0.3, 1.0, 1.0, 3.0, 6.0]), seed=r.randrange(1 << 30)))
    return out


@pytest.mark.parametrize("cfg", _fuzz_shapes(1, 40), ids=lambda c: f"T{c['T']}N{c['N']}V{c['V']}S{c['S']}x{c['scale']}")
def test_ctc_star_fuzz(hb, oracle, cfg):
    """random small shapes around the slot / warp boundaries, random lengths (0 .. S, 1 .. T), labels with
    repeats, mixed feasible and infeasible utterances, several logit scales"""
    g = torch.Generator().manual_seed(cfg["seed"])
    T, N, V, S = cfg["T"], cfg["N"], cfg["V"], cfg["S"]
    x = torch.randn(T, N, V, generator=g) * cfg["scale"]
    tg = torch.randint(1, V, (N, max(S, 1)), generator=g)[:, :S] if S else torch.zeros(N, 0, dtype=torch.long)
    il = torch.randint(1, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(0, S + 1, (N,), generator=g); tl[0] = S
    args = (tg.numpy(), il.numpy(), tl.numpy())
    for kind in ("ctc", "star"):
        if kind == "star" and (V < 2 or S > 511):
            continue
        tgz = tg.clone()
        for n in range(N):
            tgz[n, tl[n]:] = 0
        a = (tgz.numpy(),) + args[1:]
        ol, og = (oracle.ctc(x.numpy(), *a) if kind == "ctc" else oracle.star(x.numpy(), *a, star_penalty=-0.7))
        xd = x.to(dev()).requires_grad_(True)
        targs = (tgz.to(dev()), il.to(dev()), tl.to(dev()))
        loss = (hb.ctc_forward_score3(xd, *targs, from_logits=True) if kind == "ctc"
                else hb.star_ctc_forward_score(xd, *targs, star_penalty=-0.7, from_logits=True))
        fin = np.isfinite(ol)
        (loss[torch.from_numpy(fin).to(dev())]).sum().backward() if fin.any() else None
        lo = loss.detach().double().cpu().numpy()
        assert (np.isinf(lo) == np.isinf(ol)).all(), (kind, lo, ol)
        np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL, atol=1e-5, err_msg=kind)
        if fin.any():
            ogm = og.copy(); ogm[:, ~fin] = 0
            err = np.abs(xd.grad.double().cpu().numpy() - ogm).max()
            assert err < GRAD_ATOL, f"{kind} {err:.3e}"


@pytest.mark.parametrize("cfg", _fuzz_shapes(2, 24), ids=lambda c: f"T{c['T']}N{c['N']}V{c['V']}U{c['S']}x{c['scale']}")
def test_rnnt_fuzz(hb, oracle, cfg):
    g = torch.Generator().manual_seed(cfg["seed"])
    N, V, U = cfg["N"], cfg["V"], min(cfg["S"], 40)
    T = max(1, min(cfg["T"], 40))
    f = torch.randn(N, T, V, generator=g) * cfg["scale"]
    gg = torch.randn(N, U + 1, V, generator=g) * cfg["scale"]
    tg = torch.randint(0, V, (N, max(U, 1)), generator=g)[:, :U] if U else torch.zeros(N, 0, dtype=torch.long)
    il = torch.randint(1, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(0, U + 1, (N,), generator=g); tl[0] = U
    joint = (f[:, :, None, :] + gg[:, None, :, :]).contiguous()
    ol, og = oracle.rnnt(joint.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    jd = joint.to(dev()).requires_grad_(True)
    loss = hb.transducer_forward_score(jd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL, atol=1e-5)
    assert np.abs(jd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL
    fd = f.to(dev()).requires_grad_(True); gd = gg.to(dev()).requires_grad_(True)
    l2 = hb.transducer_forward_score_fg(fd, gd, tg.to(dev()), il.to(dev()), tl.to(dev()))
    l2.sum().backward()
    np.testing.assert_allclose(l2.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL, atol=1e-5)
    for got, ref in ((fd.grad, og.sum(2)), (gd.grad, og.sum(1))):
        err = np.abs(got.double().cpu().numpy() - ref)
        assert (err <= GRAD_ATOL + 2e-6 * np.abs(ref)).all(), f"{err.max():.3e}"


def test_cuda_graph_capture_and_replay(hb):
    """the ops never synchronise or allocate outside torch: a loss+gradient step can be captured once and
    replayed on new inputs written into the captured buffers (launch-bound small batches, BASELINE config 1)"""
    g = torch.Generator().manual_seed(9)
    T, N, V, S = 200, 8, 256, 50
    xs = [torch.randn(T, N, V, generator=g).to(dev()) for _ in range(3)]
    tg = torch.randint(1, V, (N, S), generator=g).to(dev())
    il = torch.full((N,), T).to(dev()); tl = torch.full((N,), S).to(dev())
    go = torch.ones(N, device=dev())
    from haloop_b200 import ops
    eager = []
    for x in xs:
        loss, ws = ops.ctc_fwd(x, tg, il, tl, True)
        eager.append((loss.clone(), ops.ctc_bwd(x, ws, go, S, True).clone()))
    xbuf = xs[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up on the capture stream
        loss, ws = ops.ctc_fwd(xbuf, tg, il, tl, True)
        ops.ctc_bwd(xbuf, ws, go, S, True)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gl, gws = ops.ctc_fwd(xbuf, tg, il, tl, True)
        gg = ops.ctc_bwd(xbuf, gws, go, S, True)
    for x, (el, eg) in zip(xs, eager):
        xbuf.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(gl, el) and torch.equal(gg, eg)      # deterministic kernels: bit-identical


# ------------------------------------------------------------- BASELINE sizes: properties ---
def _full_size_checks(kind, loss_fn, x, grad, oracle_fn, sub, row_axis_sum):
    # (1) shift invariance of a softmax-fed loss: every gradient row sums to zero
    rs = grad.sum(-1).abs().max().item()
    assert rs < 2e-5, f"{kind}: gradient rows sum to {rs:.2e}"
    # (2) a few utterances of the full batch against the float64 oracle
    for n, (ol, og) in sub.items():
        l = float(loss_fn[n])
        assert abs(l / ol - 1) < LOSS_RTOL, (kind, n, l, ol)
        err = np.abs(row_axis_sum(grad, n).double().cpu().numpy() - og).max()
        assert err < GRAD_ATOL, f"{kind} utterance {n}: {err:.3e}"


def test_full_size_ctc_c2(hb, oracle):
    """BASELINE config 2 (B=256, T=1500, V=1024, U=300): row sums, determinism, linearity in grad_output,
    and three utterances of the batch against the oracle"""
    g = torch.Generator(device=dev()).manual_seed(2)
    B, T, V, U = 256, 1500, 1024, 300
    x = torch.randn(B, T, V, device=dev(), generator=g).permute(1, 0, 2)
    tg = torch.randint(1, V, (B, U), device=dev(), generator=g)
    il = torch.randint(T // 2, T + 1, (B,), device=dev(), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (B,), device=dev(), generator=g); tl[0] = U
    from haloop_b200 import ops
    loss, ws = ops.ctc_fwd(x, tg, il, tl, True)
    g1 = ops.ctc_bwd(x, ws, torch.ones(B, device=dev()), U, True)
    g2 = ops.ctc_bwd(x, ws, torch.full((B,), 2.0, device=dev()), U, True)
    assert torch.equal(g2, 2 * g1), "gradient is linear in grad_output (exactly, for a power of two)"
    loss_b, ws_b = ops.ctc_fwd(x, tg, il, tl, True)
    assert torch.equal(loss, loss_b) and torch.equal(ops.ctc_bwd(x, ws_b, torch.ones(B, device=dev()), U, True), g1), \
        "bit-identical when repeated"
    sub = {}
    for n in (0, 17, 255):
        ol, og = oracle.ctc(x[:, n:n + 1].cpu().numpy(), tg[n:n + 1].cpu().numpy(), il[n:n + 1].cpu().numpy(),
                            tl[n:n + 1].cpu().numpy())
        sub[n] = (float(ol[0]), og[:, 0])
    _full_size_checks("ctc", loss, x, g1, None, sub, lambda gr, n: gr[:, n])
    assert not g1[int(il[17]):, 17].any(), "frames past the utterance's length carry no gradient"


def test_full_size_star_c3(hb, oracle):
    g = torch.Generator(device=dev()).manual_seed(3)
    B, T, V, U = 128, 1000, 512, 200
    x = torch.randn(B, T, V, device=dev(), generator=g).permute(1, 0, 2)
    tg = torch.randint(1, V, (B, U), device=dev(), generator=g)
    il = torch.randint(T // 2, T + 1, (B,), device=dev(), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (B,), device=dev(), generator=g); tl[0] = U
    tg = tg * (torch.arange(U, device=dev())[None, :] < tl[:, None])
    from haloop_b200 import ops
    loss, ws = ops.star_fwd(x, tg, il, tl, -0.5, True)
    g1 = ops.star_bwd(x, ws, torch.ones(B, device=dev()), U, True)
    sub = {}
    for n in (0, 77):
        ol, og = oracle.star(x[:, n:n + 1].cpu().numpy(), tg[n:n + 1].cpu().numpy(), il[n:n + 1].cpu().numpy(),
                             tl[n:n + 1].cpu().numpy(), star_penalty=-0.5)
        sub[n] = (float(ol[0]), og[:, 0])
    _full_size_checks("star", loss, x, g1, None, sub, lambda gr, n: gr[:, n])


def test_full_size_rnnt_c4(hb, oracle):
    g = torch.Generator(device=dev()).manual_seed(4)
    B, T, U, V = 32, 500, 100, 1024
    x = torch.randn(B, T, U + 1, V, device=dev(), generator=g)
    tg = torch.randint(1, V, (B, U), device=dev(), generator=g)
    il = torch.randint(T // 2, T + 1, (B,), device=dev(), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (B,), device=dev(), generator=g); tl[0] = U
    from haloop_b200 import ops
    loss, ws = ops.rnnt_fwd(x, tg, il, tl, True)
    g1 = ops.rnnt_bwd(x, ws, torch.ones(B, device=dev()), True)
    sub = {}
    for n in (0, 13):
        ol, og = oracle.rnnt(x[n:n + 1].cpu().numpy(), tg[n:n + 1].cpu().numpy(), il[n:n + 1].cpu().numpy(),
                             tl[n:n + 1].cpu().numpy())
        sub[n] = (float(ol[0]), og[0])
    _full_size_checks("rnnt", loss, x, g1, None, sub, lambda gr, n: gr[n])
    assert not g1[13, int(il[13]):].any() and not g1[13, :, int(tl[13]) + 1:].any(), "padded nodes carry no gradient"


# -------------------------------------------------------------------------------- limits ---
@pytest.mark.parametrize("cfg", [
    dict(T=3, N=65535, V=8, S=1),            # the largest batch one launch takes
    dict(T=6, N=2, V=27000, S=3),            # a vocabulary whose rows barely fit two shared-memory stages
    dict(T=30000, N=1, V=16, S=40),          # a very long utterance
    dict(T=1400, N=1, V=12, S=1023),         # the longest target
])
def test_ctc_limits(hb, oracle, cfg):
    x, tg, il, tl = _rand_ctc(600 + cfg["S"], cfg["T"], cfg["N"], cfg["V"], cfg["S"])
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    fin = np.isfinite(ol)
    loss[torch.from_numpy(fin).to(dev())].sum().backward()
    lo = loss.detach().double().cpu().numpy()
    assert (np.isinf(lo) == np.isinf(ol)).all()
    np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL)
    og[:, ~fin] = 0
    assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL


def test_rnnt_limits_and_errors(hb, oracle):
    g = torch.Generator().manual_seed(8)
    x = torch.randn(1, 3, 1024, 4, generator=g)                   # U+1 = 1024: the widest lattice
    tg = torch.randint(0, 4, (1, 1023), generator=g)
    il = torch.tensor([3]); tl = torch.tensor([1023])
    ol, og = oracle.rnnt(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.transducer_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL
    with pytest.raises(Exception):                                # unsupported shapes fail loudly, never silently
        hb.transducer_forward_score(torch.zeros(1, 2, 1026, 4, device=dev()), torch.zeros(1, 1025, dtype=torch.long),
                                    torch.tensor([2]), torch.tensor([1025]), from_logits=True)
    with pytest.raises(Exception):
        hb.ctc_forward_score3(torch.zeros(2, 70000, 4, device=dev()), torch.zeros(70000, 1, dtype=torch.long),
                              torch.full((70000,), 2), torch.ones(70000, dtype=torch.long), from_logits=True)
    with pytest.raises(Exception):
        hb.ctc_forward_score3(torch.zeros(1500, 2, 8, device=dev()), torch.ones(2, 1024, dtype=torch.long),
                              torch.full((2,), 1500), torch.full((2,), 1024), from_logits=True)


def test_custom_op_registration(hb):
    """torch.library.opcheck: schema (no hidden mutation/aliasing), fake-tensor kernels (so torch.compile traces
    through, ha/init.py:267) and autograd registration of the four forward ops"""
    from haloop_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(12, 3, 10, generator=g).to(dev()).requires_grad_(True)
    tg = torch.randint(1, 10, (3, 4), generator=g).to(dev())
    il = torch.tensor([12, 10, 9]).to(dev()); tl = torch.tensor([4, 3, 2]).to(dev())
    utils = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(ops.ctc_fwd, (x, tg, il, tl, True), test_utils=utils)
    torch.library.opcheck(ops.star_fwd, (x, tg, il, tl, -0.5, True), test_utils=utils)
    j = torch.randn(3, 6, 5, 10, generator=g).to(dev()).requires_grad_(True)
    torch.library.opcheck(ops.rnnt_fwd, (j, tg, torch.tensor([6, 5, 4]).to(dev()), tl, True), test_utils=utils)
    f = torch.randn(3, 6, 10, generator=g).to(dev()).requires_grad_(True)
    gg = torch.randn(3, 5, 10, generator=g).to(dev()).requires_grad_(True)
    torch.library.opcheck(ops.rnnt_fg_fwd, (f, gg, tg, torch.tensor([6, 5, 4]).to(dev()), tl), test_utils=utils)


def test_torch_compile_traces_through(hb):
    """--compile (ha/init.py:267): the custom ops carry fake kernels and an autograd formula, so a compiled
    training step runs them without graph breaks (aot_eager: traced forward and backward, no codegen)"""
    g = torch.Generator().manual_seed(2)
    tg = torch.randint(1, 10, (3, 4), generator=g).to(dev())
    il = torch.tensor([12, 10, 9]).to(dev()); tl = torch.tensor([4, 3, 2]).to(dev())

    def step(x):
        return hb.ctc_reduce_mean(hb.ctc_forward_score3(x * 1.5, tg, il, tl, from_logits=True), tl)

    x1 = torch.randn(12, 3, 10, generator=g).to(dev()).requires_grad_(True)
    x2 = x1.detach().clone().requires_grad_(True)
    step(x1).backward()
    compiled = torch.compile(step, backend="aot_eager", fullgraph=True)
    l2 = compiled(x2)
    l2.backward()
    assert torch.equal(step(x1.detach()), l2.detach()) and torch.equal(x1.grad, x2.grad)


def test_vmap_per_sample_gradients(hb):
    """torch.func.vmap(grad_and_value(...)) over utterances, the pattern of ha/grad_norm.py (which drops its CTC
    term because the reference has no batching rule): one merged launch, same numbers as a Python loop"""
    from torch.func import grad_and_value, vmap
    g = torch.Generator().manual_seed(12)
    B, T, V, S = 5, 30, 12, 6
    x = torch.randn(B, T, V, generator=g).to(dev())                 # (B,T,V): one utterance per vmap slice
    tg = torch.randint(1, V, (B, S), generator=g).to(dev())
    il = torch.tensor([30, 28, 20, 25, 30]).to(dev()); tl = torch.tensor([6, 5, 3, 6, 1]).to(dev())

    def one_ctc(xi, tgi, ili, tli):                                  # xi (T,V) -> scalar loss
        return hb.ctc_forward_score3(xi[:, None, :], tgi[None], ili[None], tli[None], from_logits=True)[0]

    def one_star(xi, tgi, ili, tli):
        return hb.star_ctc_forward_score(xi[:, None, :], tgi[None], ili[None], tli[None], star_penalty=-0.5,
                                         from_logits=True)[0]

    for fn in (one_ctc, one_star):
        grads, losses = vmap(grad_and_value(fn))(x, tg, il, tl)
        for b in range(B):
            gb, lb = grad_and_value(fn)(x[b], tg[b], il[b], tl[b])
            assert torch.equal(losses[b], lb) and torch.equal(grads[b], gb)

    U = 4
    f = torch.randn(B, 7, V, generator=g).to(dev()); gg = torch.randn(B, U + 1, V, generator=g).to(dev())
    tg2 = torch.randint(0, V, (B, U), generator=g).to(dev())
    il2 = torch.tensor([7, 6, 5, 7, 3]).to(dev()); tl2 = torch.tensor([4, 3, 4, 1, 2]).to(dev())

    def one_fg(fi, gi, tgi, ili, tli):
        return hb.transducer_forward_score_fg(fi[None], gi[None], tgi[None], ili[None], tli[None])[0]

    def one_joint(fi, gi, tgi, ili, tli):
        return hb.transducer_forward_score((fi[:, None, :] + gi[None, :, :])[None], tgi[None], ili[None], tli[None],
                                           from_logits=True)[0]

    for fn in (one_fg, one_joint):
        (gf, ggr), losses = vmap(grad_and_value(fn, argnums=(0, 1)))(f, gg, tg2, il2, tl2)
        for b in range(B):
            (gfb, ggb), lb = grad_and_value(fn, argnums=(0, 1))(f[b], gg[b], tg2[b], il2[b], tl2[b])
            torch.testing.assert_close(losses[b], lb, rtol=1e-6, atol=1e-6)
            torch.testing.assert_close(gf[b], gfb, rtol=1e-5, atol=2e-6)
            torch.testing.assert_close(ggr[b], ggb, rtol=1e-5, atol=2e-6)


def test_stream_ordered_on_side_streams(hb):
    """the ops launch on torch's current stream and never synchronise: two batches on two side streams, results
    equal to the default-stream ones (ha/loop.py keeps the loss on the device until loss.item())"""
    g = torch.Generator().manual_seed(21)
    outs = []
    data = []
    for k in range(2):
        x = torch.randn(300, 6, 64, generator=g).to(dev())
        tg = torch.randint(1, 64, (6, 40), generator=g).to(dev())
        il = torch.randint(150, 301, (6,), generator=g).to(dev()); tl = torch.randint(20, 41, (6,), generator=g).to(dev())
        data.append((x, tg, il, tl))
        xr = x.clone().requires_grad_(True)
        l = hb.ctc_forward_score3(xr, tg, il, tl, from_logits=True)
        l.sum().backward()
        outs.append((l.detach().clone(), xr.grad.clone()))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    res = []
    for k, st in enumerate(streams):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            x, tg, il, tl = data[k]
            xr = x.clone().requires_grad_(True)
            l = hb.ctc_forward_score3(xr, tg, il, tl, from_logits=True)
            l.sum().backward()
            res.append((l, xr))
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    for (l, xr), (l0, g0) in zip(res, outs):
        assert torch.equal(l.detach(), l0) and torch.equal(xr.grad, g0)

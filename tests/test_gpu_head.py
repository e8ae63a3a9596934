"""Fused classifier head + CTC (ha/recognizer.py:43-46 + 61-82) on the GPU.

Checked against (a) fixtures written by the unmodified reference module in float64 (tests/golden/head_temporal_
classifier.npz, adapter_temporal_classifier.npz; oracle/make_adapter_golden.py), (b) the float64 oracle
(oracle.head_ctc = Linear -> log_softmax -> CTC with the chain rule) on seeded shapes around the tile boundaries,
(c) the unfused GPU path (cuBLAS fp32 logits -> ctc_forward_score3) at BASELINE config 2 size.

Tolerances.  The op contains a D-long and an (N T)-long fp32 contraction, so gradients are compared relative to
their largest entry: "tf32x3" (the default; fp32-grade products on the tf32 tensor cores) loss 1e-5 relative,
gradients 3e-5 of max |gradient| (+1e-7 absolute); "tf32" (single product, torch's allow_tf32) loss 2e-3, gradients 3e-2 of the maximum.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_path

pytestmark = pytest.mark.gpu

LOSS_RTOL = {"tf32x3": 1e-5, "tf32": 2e-3}
GRAD_RTOL = {"tf32x3": 3e-5, "tf32": 3e-2}


@pytest.fixture(scope="module")
def hb():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import haloop_b200
    return haloop_b200


def _close(name, got, want, rtol):
    got = got.detach().double().cpu().numpy()
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= rtol * scale + 1e-7, f"{name}: error {err:.3e} vs max |gradient| {scale:.3e} (allowed {rtol:.0e} of it)"


def _run(hb, h, W, b, tg, il, tl, go, precision="tf32x3"):
    hh = torch.tensor(h, dtype=torch.float32, device="cuda").requires_grad_(True)
    WW = torch.tensor(W, dtype=torch.float32, device="cuda").requires_grad_(True)
    bb = None if b is None else torch.tensor(b, dtype=torch.float32, device="cuda").requires_grad_(True)
    loss = hb.linear_ctc_forward_score(hh, WW, bb, torch.as_tensor(tg).cuda(), torch.as_tensor(il).cuda(),
                                       torch.as_tensor(tl).cuda(), precision=precision)
    (loss * torch.tensor(go, dtype=torch.float32, device="cuda")).sum().backward()
    return loss, hh.grad, WW.grad, None if bb is None else bb.grad


def _check(hb, oracle, h, W, b, tg, il, tl, go, precision="tf32x3"):
    loss, dh, dW, db = _run(hb, h, W, b, tg, il, tl, go, precision)
    ol, odh, odW, odb = oracle.head_ctc(np.float32(h), np.float32(W), None if b is None else np.float32(b), tg, il, tl, go)
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    assert (np.isfinite(lo) == fin).all()
    assert np.abs(lo[fin] / ol[fin] - 1).max() < LOSS_RTOL[precision]
    _close("dh", dh, odh, GRAD_RTOL[precision])
    _close("dW", dW, odW, GRAD_RTOL[precision])
    if b is not None:
        _close("db", db, odb, GRAD_RTOL[precision])


def test_head_vs_reference_module_golden(hb):
    """TemporalClassifier.log_probs + ctc_forward_score3 of the unmodified reference, float64."""
    d = np.load(golden_path("head_temporal_classifier"))
    loss, dh, dW, db = _run(hb, d["feats"], d["weight"], d["bias"], d["targets"], d["in_len"], d["tgt_len"], d["grad_out"])
    assert np.abs(loss.detach().double().cpu().numpy() / d["loss"] - 1).max() < 1e-5
    _close("dh", dh, d["dh"], 3e-5)
    _close("dW", dW, d["dW"], 3e-5)
    _close("db", db, d["db"], 3e-5)


def test_head_mean_reduction_vs_live_call_site_golden(hb):
    """The live call site (ha/recognizer.py:71, F.ctc_loss 'mean') with an empty transcript in the batch; V = 9
    (not a multiple of 4)."""
    d = np.load(golden_path("adapter_temporal_classifier"))
    W = torch.tensor(d["weight"], dtype=torch.float32, device="cuda").requires_grad_(True)
    b = torch.tensor(d["bias"], dtype=torch.float32, device="cuda").requires_grad_(True)
    h = torch.tensor(d["feats"], dtype=torch.float32, device="cuda")
    loss = hb.linear_ctc_loss(h, W, b, torch.as_tensor(d["targets"]).cuda(), torch.as_tensor(d["in_len"]).cuda(),
                              torch.as_tensor(d["tgt_len"]).cuda())
    loss.backward()
    assert abs(float(loss) / float(d["ctc_loss"]) - 1) < 1e-5
    _close("dW", W.grad, d["ctc_gw"], 3e-5)
    _close("db", b.grad, d["ctc_gb"], 3e-5)


def _case(seed, N, T, D, V, S, repeats=False, scale=1.0):
    g = np.random.default_rng(seed)
    h = g.standard_normal((N, T, D)).astype(np.float32)
    W = (g.standard_normal((V, D)) * scale / np.sqrt(D)).astype(np.float32)
    b = (g.standard_normal(V) * 0.2).astype(np.float32)
    tg = g.integers(1, 3 if repeats else V, (N, S))
    il = g.integers(max(T // 2, min(T, 2 * S + 1)), T + 1, N); il[0] = T
    tl = g.integers(max(S // 2, 1), S + 1, N); tl[0] = S
    go = g.uniform(0.5, 1.5, N)
    return h, W, b, tg, il, tl, go


@pytest.mark.parametrize("N,T,D,V,S", [
    (3, 70, 64, 40, 7),            # one tile, everything ragged
    (2, 130, 100, 260, 11),        # D not a multiple of the 32-wide k block, V = 2 tiles + 4 classes
    (5, 127, 32, 128, 30),         # rows one short of a tile multiple
    (4, 300, 256, 1000, 60),       # 8 class tiles, the last one ragged
    (1, 40, 8, 5, 3),              # tiny: V < 16, D < 32
    (2, 64, 36, 1026, 9),          # class count not a multiple of 4, 9 tiles
])
def test_head_vs_oracle_shapes(hb, oracle, N, T, D, V, S):
    _check(hb, oracle, *_case(N * 1000 + V, N, T, D, V, S))


def test_head_repeated_labels_and_peaky_logits(hb, oracle):
    _check(hb, oracle, *_case(7, 4, 120, 64, 48, 25, repeats=True))
    _check(hb, oracle, *_case(8, 3, 200, 128, 96, 20, scale=6.0))


def test_head_without_bias_and_single_product_mode(hb, oracle):
    h, W, b, tg, il, tl, go = _case(11, 3, 90, 96, 200, 12)
    _check(hb, oracle, h, W, None, tg, il, tl, go)
    _check(hb, oracle, h, W, b, tg, il, tl, go, precision="tf32")


def test_head_infeasible_and_empty_utterances(hb, oracle):
    """T_n < L_n (+inf loss, zero gradient, as the oracle and F.ctc_loss give) and an empty transcript."""
    h, W, b, tg, il, tl, go = _case(13, 4, 60, 64, 50, 20)
    il[1] = 10; tl[1] = 20           # infeasible
    tl[2] = 0                        # empty transcript: all blanks
    loss, dh, dW, db = _run(hb, h, W, b, tg, il, tl, go)
    ol, odh, odW, odb = oracle.head_ctc(h, W, b, tg, il, tl, go)
    lo = loss.detach().double().cpu().numpy()
    assert np.isinf(lo[1]) and np.isinf(ol[1])
    fin = np.isfinite(ol)
    assert np.abs(lo[fin] / ol[fin] - 1).max() < 1e-5
    assert float(dh[1].abs().max()) == 0.0
    _close("dh", dh, odh, 3e-5); _close("dW", dW, odW, 3e-5); _close("db", db, odb, 3e-5)


def test_head_backward_in_several_row_chunks(hb, oracle):
    """N T = 12000 rows with V = D = 1024: the backward runs in two row chunks (8192 + 3808) and two split-K halves."""
    _check(hb, oracle, *_case(17, 12, 1000, 1024, 1024, 50))


def test_head_matches_the_unfused_path_at_config2_size(hb):
    """BASELINE config 2 (B=256, T=1500, V=1024, U=300) with the reference's feat_dim D=1024 (ha/recognizer.py:38):
    against cuBLAS fp32 logits -> ctc_forward_score3 -> autograd on the same GPU (that path is oracle-checked at this
    size in test_gpu_round2.py).  Also: the bias gradient sums to zero (softmax - occupancy rows do)."""
    B, T, D, V, U = 256, 1500, 1024, 1024, 300
    g = torch.Generator(device="cuda").manual_seed(5)
    h = torch.randn(B, T, D, device="cuda", generator=g)
    W = (torch.randn(V, D, device="cuda", generator=g) / D ** 0.5).requires_grad_(True)
    b = (torch.randn(V, device="cuda", generator=g) * 0.1).requires_grad_(True)
    tg = torch.randint(1, V, (B, U), device="cuda", generator=g)
    il = torch.randint(T // 2, T + 1, (B,), device="cuda", generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (B,), device="cuda", generator=g); tl[0] = U
    h.requires_grad_(True)
    loss = hb.linear_ctc_forward_score(h, W, b, tg, il, tl)
    loss.sum().backward()
    fused = [loss.detach().clone(), h.grad.clone(), W.grad.clone(), b.grad.clone()]
    h.grad = W.grad = b.grad = None
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = hb.ctc_forward_score3(F.linear(h, W, b).permute(1, 0, 2), tg, il, tl, from_logits=True)
        ref.sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert float(((fused[0] - ref) / ref).abs().max()) < 1e-5
    for name, a, r in (("dh", fused[1], h.grad), ("dW", fused[2], W.grad), ("db", fused[3], b.grad)):
        err, scale = float((a - r).abs().max()), float(r.abs().max())
        assert err <= 5e-5 * scale, f"{name}: {err:.3e} vs max {scale:.3e}"
    assert abs(float(fused[3].double().sum())) < 1e-3 * float(fused[3].abs().max())


def test_head_rejects_what_it_cannot_do(hb):
    h = torch.randn(2, 10, 6, device="cuda")             # D not a multiple of 4
    W = torch.randn(8, 6, device="cuda")
    tg = torch.ones(2, 2, dtype=torch.long, device="cuda")
    il = torch.tensor([10, 10], device="cuda"); tl = torch.tensor([2, 2], device="cuda")
    with pytest.raises(Exception):
        hb.linear_ctc_forward_score(h, W, None, tg, il, tl)
    with pytest.raises(ValueError):
        hb.linear_ctc_forward_score(h.cpu(), W.cpu(), None, tg.cpu(), il.cpu(), tl.cpu())
    with pytest.raises(ValueError):
        hb.linear_ctc_forward_score(torch.randn(2, 10, 8, device="cuda"), torch.randn(8, 8, device="cuda"), None, tg, il, tl,
                                    precision="bf16")


def test_head_custom_op_registration_and_double_backward(hb):
    """torch.library.opcheck on the forward op (schema, fake kernel, autograd registration) and a loud failure for
    second-order use (the backward op has no derivative)."""
    from haloop_b200 import ops
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 20, 16, generator=g).cuda().requires_grad_(True)
    W = torch.randn(12, 16, generator=g).cuda().requires_grad_(True)
    b = torch.randn(12, generator=g).cuda().requires_grad_(True)
    tg = torch.randint(1, 12, (2, 4), generator=g).cuda()
    il = torch.tensor([20, 15]).cuda(); tl = torch.tensor([4, 2]).cuda()
    torch.library.opcheck(ops.head_ctc_fwd, (h, W, b, tg, il, tl, 3),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    loss = hb.linear_ctc_forward_score(h, W, b, tg, il, tl).sum()
    (gh,) = torch.autograd.grad(loss, h, create_graph=True)      # first order works; the op is once-differentiable:
    with pytest.raises((NotImplementedError, RuntimeError)):
        gh.sum().backward()


def test_head_cuda_graph_capture_and_replay(hb):
    """The op builds its tensor maps on the host and launches on the current stream without synchronising: a forward +
    backward pair captures into a CUDA graph and replays bit-identically on new features written into the captured
    buffer."""
    from haloop_b200 import ops
    g = torch.Generator().manual_seed(4)
    N, T, D, V, S = 4, 200, 128, 256, 30
    hs = [torch.randn(N, T, D, generator=g).cuda() for _ in range(3)]
    W = (torch.randn(V, D, generator=g) / D ** 0.5).cuda(); b = torch.randn(V, generator=g).cuda()
    tg = torch.randint(1, V, (N, S), generator=g).cuda()
    il = torch.tensor([200, 180, 150, 200]).cuda(); tl = torch.tensor([30, 25, 10, 30]).cuda()
    go = torch.ones(N, device="cuda")
    eager = []
    for h in hs:
        loss, saved = ops.head_ctc_fwd(h, W, b, tg, il, tl, 3)
        eager.append((loss.clone(),) + tuple(t.clone() for t in ops.head_ctc_bwd(h, W, b, saved, go, S, 3)))
    hbuf = hs[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss, saved = ops.head_ctc_fwd(hbuf, W, b, tg, il, tl, 3)
        ops.head_ctc_bwd(hbuf, W, b, saved, go, S, 3)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gl, gsaved = ops.head_ctc_fwd(hbuf, W, b, tg, il, tl, 3)
        gdh, gdW, gdb = ops.head_ctc_bwd(hbuf, W, b, gsaved, go, S, 3)
    for h, (el, edh, edW, edb) in zip(hs, eager):
        hbuf.copy_(h)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(gl, el) and torch.equal(gdh, edh) and torch.equal(gdW, edW) and torch.equal(gdb, edb)

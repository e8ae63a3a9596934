"""The fused star-CTC path (csrc/star2.cuh: two kernels, emissions and occupancies never in HBM) against the float64
oracle (oracle/ha_oracle.c restating ha/star.py:65-163) on the cases the reference's semantics make special:
labels stored beyond L_n (the last star reads targets[n, L_n], ha/star.py:46), L_n == S (the all-star), label 0 inside a
transcript, repeated labels, the log-prob boundary, several penalties, T_n = 1, L_n = 0, infeasible utterances, logits
outside the one-pass range of the row statistics.  Tolerances: loss 1e-4 relative, gradient 1e-5 absolute.
"""
import numpy as np
import pytest
import torch

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


def _run(hb, oracle, x, tg, il, tl, pen, from_logits=True, go=None):
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=pen, from_logits=from_logits,
                         grad_out=None if go is None else go.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), star_penalty=pen, from_logits=from_logits)
    (loss if go is None else loss * go.to(dev())).sum().backward()
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    assert (np.isinf(lo) == np.isinf(ol)).all(), (lo, ol)
    np.testing.assert_allclose(lo[fin], ol[fin], rtol=LOSS_RTOL)
    g = xd.grad.double().cpu().numpy()
    assert np.isfinite(g).all()
    err = np.abs(g - og).max()
    assert err < GRAD_ATOL, f"{err:.3e}"
    return lo, g


def test_fused_path_is_the_one_that_runs(hb):
    from haloop_b200 import _lib
    L = _lib.lib()
    # the fused layout has no emission rows: far smaller than the three-kernel path's workspace for the same shape
    fused = L.ha_star_workspace_bytes(1000, 128, 512, 200)
    legacy = L.ha_star_workspace_bytes(1000, 128, 511, 200)        # V % 4 != 0: three-kernel path
    assert fused < legacy


@pytest.mark.parametrize("cfg", [
    dict(T=120, N=6, V=32, S=20, pen=-0.5),
    dict(T=200, N=4, V=64, S=70, pen=0.0),                   # 2 trellis warps
    dict(T=300, N=3, V=16, S=130, pen=-3.0),                 # 2 warps
    dict(T=400, N=2, V=8, S=300, pen=-0.25),                 # 3 warps
    dict(T=700, N=2, V=12, S=500, pen=-0.5),                 # 5 warps
    dict(T=64, N=8, V=512, S=9, pen=-1.0),
])
def test_star2_ragged_vs_oracle(hb, oracle, cfg):
    g = torch.Generator().manual_seed(cfg["T"] + cfg["S"])
    T, N, V, S = cfg["T"], cfg["N"], cfg["V"], cfg["S"]
    x = torch.randn(T, N, V, generator=g) * 1.5
    tg = torch.randint(1, V, (N, S), generator=g)            # labels beyond L_n stay in place: ha/star.py:46 reads them
    il = torch.randint(T // 2 + S // 2, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(0, S + 1, (N,), generator=g); tl[0] = S; tl[-1] = 0
    go = torch.linspace(0.5, 2.0, N)
    _run(hb, oracle, x, tg, il, tl, cfg["pen"], go=go)


@pytest.mark.parametrize("cfg", [
    dict(T=50, N=3, V=4096, S=20),       # rows of 16 KB: 4 row warps, two ring stages
    dict(T=40, N=2, V=8192, S=20),       # rows of 32 KB: one stage per row warp
    dict(T=30, N=3, V=4, S=6),           # the smallest class count of the fused path
])
def test_star2_row_ring_configurations(hb, oracle, cfg):
    g = torch.Generator().manual_seed(cfg["V"])
    T, N, V, S = cfg["T"], cfg["N"], cfg["V"], cfg["S"]
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.randint(T // 2 + S, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(1, S + 1, (N,), generator=g); tl[0] = S
    _run(hb, oracle, x, tg, il, tl, -0.5)


def test_star2_empty_target_dimension(hb, oracle):
    """S = 0: every utterance is the all-star .* (ha/star.py:47) around blanks"""
    g = torch.Generator().manual_seed(3)
    T, N, V = 25, 3, 16
    x = torch.randn(T, N, V, generator=g)
    tg = torch.zeros(N, 0, dtype=torch.long)
    il = torch.tensor([25, 1, 12]); tl = torch.zeros(N, dtype=torch.long)
    _run(hb, oracle, x, tg, il, tl, -0.5)


def test_star2_unaligned_and_permuted_views(hb, oracle):
    """a class-sliced view (rows not 16-byte aligned: copied once by the shim, as for CTC) and the (N,T,C) buffer the
    reference hands the loss as a permuted (T,N,C) view (taken as it is, gradient with the same strides)"""
    g = torch.Generator().manual_seed(11)
    T, N, V, S = 40, 3, 16, 6
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.tensor([40, 33, 21]); tl = torch.tensor([6, 4, 2])
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5)
    wide = torch.zeros(T, N, V + 2, device=dev()); wide[..., 1:V + 1] = x.to(dev())
    ntc = x.permute(1, 0, 2).contiguous().to(dev())
    for view in (wide[..., 1:V + 1], ntc.permute(1, 0, 2)):
        xd = view.detach().requires_grad_(True)
        loss = hb.star_ctc_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), star_penalty=-0.5, from_logits=True)
        loss.sum().backward()
        np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
        assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL


def test_star2_label_zero_and_repeats(hb, oracle):
    g = torch.Generator().manual_seed(5)
    T, N, V, S = 150, 6, 8, 40
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(0, 3, (N, S), generator=g)            # label 0 and long runs of equal labels, > 8 occurrences a class
    il = torch.full((N,), T); tl = torch.randint(S // 2, S + 1, (N,), generator=g); tl[0] = S
    _run(hb, oracle, x, tg, il, tl, -0.5)


def test_star2_log_prob_boundary(hb, oracle):
    g = torch.Generator().manual_seed(6)
    T, N, V, S = 90, 4, 24, 12
    lp = torch.randn(T, N, V, generator=g).log_softmax(-1)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.tensor([90, 60, 33, 90]); tl = torch.tensor([12, 7, 0, 3])
    _run(hb, oracle, lp, tg, il, tl, -0.5, from_logits=False)


def test_star2_tiny_and_infeasible(hb, oracle):
    g = torch.Generator().manual_seed(7)
    T, N, V, S = 12, 6, 8, 6
    x = torch.randn(T, N, V, generator=g)
    tg = torch.tensor([[1, 1, 1, 1, 1, 1], [1, 2, 3, 4, 5, 6], [2, 2, 3, 3, 4, 4], [1, 2, 1, 2, 1, 2], [3, 3, 3, 1, 1, 1], [5, 5, 5, 5, 5, 5]])
    il = torch.tensor([12, 6, 8, 1, 11, 10]); tl = torch.tensor([6, 6, 6, 1, 6, 6])      # 11 frames exactly fit 6 labels + 5 gaps; 10 do not
    lo, gr = _run(hb, oracle, x, tg, il, tl, -0.5)
    assert np.isinf(lo[5]) and np.isinf(lo[2]) and np.isfinite(lo[0]) and np.isfinite(lo[4])
    assert (gr[:, 5] == 0).all() and (gr[1:, 3] == 0).all()


@pytest.mark.parametrize("shift", [120.0, -150.0])
def test_star2_rows_outside_the_one_pass_range(hb, oracle, shift):
    """rows whose plain exponential sums leave the fp32 range take the shifted two-pass statistics"""
    g = torch.Generator().manual_seed(8)
    T, N, V, S = 60, 3, 16, 8
    x = torch.randn(T, N, V, generator=g)
    x[::3] += shift
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.full((N,), T); tl = torch.tensor([8, 5, 2])
    xd = x.to(dev()).requires_grad_(True)
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), star_penalty=-0.5, from_logits=True)
    loss.sum().backward()
    np.testing.assert_allclose(loss.detach().double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    # a logit of 150 carries an fp32 ulp of 1.5e-5 nats: the softmax term is good to |x|max * 2^-23 relative
    assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL * abs(shift) / 32.0


def test_star2_bit_identical_repeats_and_zero_rows(hb):
    g = torch.Generator().manual_seed(9)
    T, N, V, S = 257, 16, 128, 60
    x = torch.randn(T, N, V, generator=g).to(dev())
    tg = torch.randint(1, V, (N, S), generator=g).to(dev())
    il = torch.randint(T // 2, T + 1, (N,), generator=g).to(dev())
    tl = torch.randint(1, S + 1, (N,), generator=g).to(dev())
    from haloop_b200 import ops
    outs = []
    for _ in range(3):
        loss, ws = ops.star_fwd(x, tg, il, tl, -0.5, True)
        gr = ops.star_bwd(x, ws, torch.ones(N, device=dev()), S, True)
        outs.append((loss.clone(), gr.clone()))
    for l2, g2 in outs[1:]:
        assert torch.equal(outs[0][0], l2) and torch.equal(outs[0][1], g2)
    gr = outs[0][1]
    for n in range(N):
        assert (gr[int(il[n]):, n] == 0).all()
    assert gr.sum(-1).abs().max() < 2e-5, "gradient rows sum to zero through the fused log-softmax"


def test_star2_cuda_graph_capture_and_replay(hb):
    """the fused star-CTC ops neither synchronise nor allocate outside torch: a loss + gradient step is captured once and
    replayed on new inputs, bit-identical to the eager calls"""
    g = torch.Generator().manual_seed(10)
    T, N, V, S = 120, 6, 64, 30
    xs = [torch.randn(T, N, V, generator=g).to(dev()) for _ in range(3)]
    tg = torch.randint(1, V, (N, S), generator=g).to(dev())
    il = torch.full((N,), T).to(dev()); tl = torch.randint(1, S + 1, (N,), generator=g).to(dev())
    go = torch.ones(N, device=dev())
    from haloop_b200 import ops
    eager = []
    for x in xs:
        loss, ws = ops.star_fwd(x, tg, il, tl, -0.5, True)
        eager.append((loss.clone(), ops.star_bwd(x, ws, go, S, True).clone()))
    xbuf = xs[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss, ws = ops.star_fwd(xbuf, tg, il, tl, -0.5, True)
        ops.star_bwd(xbuf, ws, go, S, True)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gl, gws = ops.star_fwd(xbuf, tg, il, tl, -0.5, True)
        gg = ops.star_bwd(xbuf, gws, go, S, True)
    for x, (el, eg) in zip(xs, eager):
        xbuf.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(gl, el) and torch.equal(gg, eg)

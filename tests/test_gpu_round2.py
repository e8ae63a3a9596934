"""Round-2 parity tests (VERDICT r01 "parity gaps"): every utterance of the BASELINE batches against the float64
oracle, x3 logits at full size, the emission floor with 100-300 nat logit gaps, the length-bucketed sweep path,
the loss all-reduce under NCCL, nested greedy decoding, the adapter with an empty transcript, double backward.

Tolerances (BASELINE.json north_star): loss 1e-4 relative, logit gradient 1e-5 absolute, against float64 gold.
"""
import os
import socket

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


def _batch(seed, B, T, V, U, scale=1.0, joint=False):
    g = torch.Generator(device=dev()).manual_seed(seed)
    shape = (B, T, U + 1, V) if joint else (B, T, V)
    x = torch.randn(shape, device=dev(), generator=g) * scale
    tg = torch.randint(1, V, (B, U), device=dev(), generator=g)
    il = torch.randint(T // 2, T + 1, (B,), device=dev(), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (B,), device=dev(), generator=g); tl[0] = U
    return x, tg, il, tl


def _check_all(kind, loss, grad, ol, og, batch_axis):
    lo = loss.double().cpu().numpy()
    fin = np.isfinite(ol)
    assert (np.isfinite(lo) == fin).all()
    rel = np.abs(lo[fin] / ol[fin] - 1).max()
    assert rel < LOSS_RTOL, f"{kind}: loss rel err {rel:.2e}"
    g = grad.double().cpu().numpy()
    err = np.abs(g - og).reshape(-1).max() if batch_axis is None else np.abs(g - og).max()
    assert err < GRAD_ATOL, f"{kind}: worst gradient error over the whole batch {err:.3e}"


# ------------------------------------------------- every utterance of the BASELINE batches ---
@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_ctc_c2_every_utterance(hb, oracle, scale):
    """BASELINE config 2 (B=256, T=1500, V=1024, U=300), logits x1 and x3 (peaky posteriors): all 256 losses and
    the full (T,B,V) gradient against the oracle."""
    from haloop_b200 import ops
    B, T, V, U = 256, 1500, 1024, 300
    x, tg, il, tl = _batch(20 + int(scale), B, T, V, U, scale)
    xv = x.permute(1, 0, 2)
    loss, ws = ops.ctc_fwd(xv, tg, il, tl, True)
    g = ops.ctc_bwd(xv, ws, torch.ones(B, device=dev()), U, True)
    assert ws.numel() < 0.65 * x.numel() * 4, "state saved for backward: stored label states only"
    ol, og = oracle.ctc(xv.cpu().numpy(), tg.cpu().numpy(), il.cpu().numpy(), tl.cpu().numpy())
    _check_all("ctc", loss, g, ol, og, 1)


@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_star_c3_every_utterance(hb, oracle, scale):
    from haloop_b200 import ops
    B, T, V, U = 128, 1000, 512, 200
    x, tg, il, tl = _batch(30 + int(scale), B, T, V, U, scale)
    tg = tg * (torch.arange(U, device=dev())[None, :] < tl[:, None])
    xv = x.permute(1, 0, 2)
    loss, ws = ops.star_fwd(xv, tg, il, tl, -0.5, True)
    g = ops.star_bwd(xv, ws, torch.ones(B, device=dev()), U, True)
    ol, og = oracle.star(xv.cpu().numpy(), tg.cpu().numpy(), il.cpu().numpy(), tl.cpu().numpy(), star_penalty=-0.5)
    _check_all("star", loss, g, ol, og, 1)


@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_rnnt_c4_eight_utterances(hb, oracle, scale):
    """BASELINE config 4 shape (T=500, U=100, V=1024): the kernels run the full batch of 32; 8 utterances (0.4 GB of
    float64 joint each on the host) are compared with the oracle."""
    from haloop_b200 import ops
    B, T, U, V = 32, 500, 100, 1024
    x, tg, il, tl = _batch(40 + int(scale), B, T, V, U, scale, joint=True)
    loss, ws = ops.rnnt_fwd(x, tg, il, tl, True)
    g = ops.rnnt_bwd(x, ws, torch.ones(B, device=dev()), True)
    pick = [0, 3, 7, 12, 13, 21, 30, 31]
    for n in pick:
        ol, og = oracle.rnnt(x[n:n + 1].cpu().numpy(), tg[n:n + 1].cpu().numpy(), il[n:n + 1].cpu().numpy(),
                             tl[n:n + 1].cpu().numpy())
        assert abs(float(loss[n]) / ol[0] - 1) < LOSS_RTOL
        err = np.abs(g[n].double().cpu().numpy() - og[0]).max()
        assert err < GRAD_ATOL, f"rnnt utterance {n}: {err:.3e}"


# ------------------------------------------------------------------- the emission floor ---
FLOOR_NATS = 87.0          # 2^-125.75 (fused CTC path) / 2^-126 below the row maximum (star, RNN-T): DESIGN.md, Limits


@pytest.mark.parametrize("gap", [100.0, 200.0, 300.0])
def test_emission_floor_is_harmless_when_the_path_can_take_the_likely_class(hb, oracle, gap):
    """Frames in which one class is 100-300 nats above all others: as long as the alignment may emit that class
    (here the blank) nothing it needs sits on the floor, and loss and gradient match the oracle as usual."""
    g = torch.Generator().manual_seed(int(gap))
    T, N, V, S = 80, 4, 32, 10
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.full((N,), T); tl = torch.full((N,), S)
    hot = torch.randperm(T, generator=g)[:12]
    x[hot, :, 0] += gap                                   # the blank dominates these frames
    for name, fn, ofn in (
        ("ctc", lambda xx: hb.ctc_forward_score3(xx, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True),
         lambda: oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())),
        ("star", lambda xx: hb.star_ctc_forward_score(xx, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True),
         lambda: oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5)),
    ):
        xd = x.to(dev()).requires_grad_(True)
        loss = fn(xd)
        loss.sum().backward()
        ol, og = ofn()
        assert np.abs(loss.detach().double().cpu().numpy() / ol - 1).max() < LOSS_RTOL, name
        # a logit of `gap` nats carries an fp32 ulp of up to 3e-5 nats, and so does the row log-sum-exp the softmax is
        # formed from: the dense term of the gradient is good to |x|max * 2^-23 relative (Limits)
        assert np.abs(xd.grad.double().cpu().numpy() - og).max() < GRAD_ATOL * max(1.0, gap / 64.0), name


@pytest.mark.parametrize("gap", [100.0, 200.0, 300.0])
def test_emission_floor_bounds_the_deviation_when_the_path_is_forced_through_it(hb, oracle, gap):
    """The documented deviation from the reference (which has no floor in its log domain): a class that is neither
    blank nor in the transcript dominates `k` frames by `gap` nats, so every alignment must emit something `gap` nats
    down in those frames.  The reference charges `gap` per frame, the kernels charge at most the floor (87 nats):
    the loss is finite, never above the oracle's, and short of it by at most k * (gap - floor)."""
    g = torch.Generator().manual_seed(7 + int(gap))
    T, N, V, S = 60, 3, 24, 8
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V - 1, (N, S), generator=g)      # class V-1 never appears in a transcript
    il = torch.full((N,), T); tl = torch.full((N,), S)
    k = 5
    hot = torch.randperm(T, generator=g)[:k]
    x[hot, :, V - 1] += gap
    ol, _ = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    loss.sum().backward()
    lo = loss.detach().double().cpu().numpy()
    assert np.isfinite(lo).all() and torch.isfinite(xd.grad).all()
    assert (lo <= ol * (1 + LOSS_RTOL)).all(), "the floor can only make a forced path cheaper"
    assert (lo >= ol - k * (gap - FLOOR_NATS) - 6.0 * k).all(), (lo, ol)        # 6 nats: the randn part of a logit gap
    assert xd.grad.sum(-1).abs().max() < 2e-5, "gradient rows still sum to zero"


# ------------------------------------------------------------ the length-bucketed sweep ---
@pytest.mark.parametrize("kind", ["ctc", "rnnt"])
def test_sweep_path_matches_the_oracle_per_utterance(hb, oracle, kind):
    """bench.py's config-5 step in miniature: deal a pool of variable-length utterances to 2 ranks by cost, cut each
    share into length buckets, run the buckets over 3 streams (sharding.bucketed_pass) and compare EVERY utterance's
    loss and gradient with the oracle."""
    import random
    from haloop_b200 import sharding
    rnd = random.Random(3)
    V, n_pool = 64, 48
    if kind == "ctc":
        tl_ = [4 * rnd.randint(10, 60) for _ in range(n_pool)]
        ul_ = [max(1, min((t - 1) // 2, round(t / 5 * rnd.uniform(0.6, 1.0)))) for t in tl_]
        budget = 12 * 240 * V * 4
    else:
        tl_ = [rnd.randint(10, 40) for _ in range(n_pool)]
        ul_ = [rnd.randint(2, 12) for _ in range(n_pool)]
        budget = 6 * 40 * 13 * V * 4
    shares = sharding.deal_utterances(tl_, ul_, V, 2, kind)
    assert sorted(shares[0] + shares[1]) == list(range(n_pool))
    g = torch.Generator().manual_seed(11)
    seen = 0
    streams = [torch.cuda.Stream() for _ in range(3)]
    for share in shares:
        buckets = sharding.bucket_by_length([tl_[i] for i in share], [ul_[i] for i in share], V, budget, kind)
        assert len(buckets) >= 2
        data = []
        for b in buckets:
            ids = [share[i] for i in b.indices]
            Bk = len(ids)
            shape = (Bk, b.t_max, V) if kind == "ctc" else (Bk, b.t_max, b.u_max + 1, V)
            x = torch.randn(shape, generator=g)
            il = torch.tensor([tl_[i] for i in ids]); tl = torch.tensor([ul_[i] for i in ids])
            tg = torch.randint(1, V, (Bk, b.u_max), generator=g)
            tg = tg * (torch.arange(b.u_max)[None, :] < tl[:, None])
            data.append((x, tg, il, tl))
        res = sharding.bucketed_pass(kind, [(x.to(dev()), tg.to(dev()), il.to(dev()), tl.to(dev()),
                                             torch.ones(len(il), device=dev())) for x, tg, il, tl in data], streams)
        torch.cuda.synchronize()
        for (x, tg, il, tl), (loss, grad) in zip(data, res):
            if kind == "ctc":
                ol, og = oracle.ctc(x.permute(1, 0, 2).numpy(), tg.numpy(), il.numpy(), tl.numpy())
                og = og.transpose(1, 0, 2)
            else:
                ol, og = oracle.rnnt(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
            assert np.abs(loss.double().cpu().numpy() / ol - 1).max() < LOSS_RTOL
            assert np.abs(grad.double().cpu().numpy() - og).max() < GRAD_ATOL
            seen += len(il)
    assert seen == n_pool


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    from haloop_b200 import sharding
    import haloop_b200 as hb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=d)
    try:
        g = torch.Generator().manual_seed(5)
        T, N, V, S = 50, 6, 16, 7
        x = torch.randn(T, N, V, generator=g); tg = torch.randint(1, V, (N, S), generator=g)
        il = torch.full((N,), T); tl = torch.randint(3, S + 1, (N,), generator=g)
        mine = sharding.shard_batch(N, rank, world)
        sl = slice(mine.start, mine.stop)
        xd = x[:, sl].contiguous().to(d).requires_grad_(True)
        loss = hb.ctc_forward_score3(xd, tg[sl].to(d), il[sl].to(d), tl[sl].to(d), from_logits=True)
        tot = sharding.reduce_loss(loss, 1.0 / tl[sl].to(d).float())
        tot.backward()
        out[rank] = (float(tot), xd.grad.cpu().numpy(), (mine.start, mine.stop))
    finally:
        dist.destroy_process_group()


def test_reduce_loss_under_nccl_two_gpus(hb, oracle):
    """The path's one collective on real GPUs: two ranks, each with half of a batch; every rank gets the global
    ctc_reduce_mean (ha/ctc.py:177-178) and the gradient of its own utterances."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    T, N, V, S = 50, 6, 16, 7
    x = torch.randn(T, N, V, generator=g); tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.full((N,), T); tl = torch.randint(3, S + 1, (N,), generator=g)
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), grad_out=(1.0 / tl.double().numpy()) / N)
    ref = oracle.ctc_reduce_mean(ol, tl.numpy())
    for r in range(world):
        tot, grad, (lo, hi) = out[r]
        assert abs(tot / ref - 1) < LOSS_RTOL
        assert np.abs(grad - og[:, lo:hi]).max() < GRAD_ATOL


# ------------------------------------------------------------------------------ decoding ---
def test_greedy_decode_nested_matches_the_reference_semantics(hb):
    """TemporalClassifier.decode (ha/recognizer.py:48-59): per-frame argmax, collapse repeats, drop blanks; honouring
    input_lengths, as a nested tensor."""
    g = torch.Generator().manual_seed(9)
    N, T, V = 5, 70, 12
    lp = torch.randn(N, T, V, generator=g).log_softmax(-1)
    lp[:, 1::2] = lp[:, ::2]                                      # force repeats
    il = torch.tensor([70, 1, 33, 64, 70])
    from haloop_b200 import align
    hyps = align.greedy_decode_nested(lp.to(dev()), il.to(dev()))[0]
    parts = hyps.unbind() if hasattr(hyps, "unbind") else list(hyps)
    assert len(parts) == N
    for n in range(N):
        a = lp[n, :int(il[n])].max(dim=-1).indices
        ref = torch.unique_consecutive(a)
        ref = ref[ref != 0]
        assert torch.equal(parts[n].cpu().long(), ref), n


def test_viterbi_of_an_infeasible_utterance_is_all_minus_one(hb):
    from haloop_b200 import ops
    g = torch.Generator().manual_seed(2)
    T, N, V = 6, 2, 8
    lp = torch.randn(T, N, V, generator=g).log_softmax(-1).to(dev())
    tg = torch.tensor([[1, 1, 1, 1], [2, 3, 0, 0]], device=dev())    # [1,1,1,1] needs 7 frames, only 6 given
    il = torch.tensor([6, 6], device=dev()); tl = torch.tensor([4, 2], device=dev())
    ali, sc = ops.ctc_viterbi(lp, tg, il, tl)
    assert torch.isinf(sc[0]) and sc[0] < 0 and (ali[0] == -1).all()
    assert torch.isfinite(sc[1]) and (ali[1] >= 0).all()


# ----------------------------------------------------------------------------- ADVICE r01 ---
def test_label_zero_in_targets_needs_its_blank(hb, oracle):
    """A label equal to 0 can only be entered from the blank before it (ha/ctc.py:140), so [3, 0] needs 3 frames:
    with 2 the utterance is infeasible (+inf, zero gradient), with 3 it matches the oracle."""
    g = torch.Generator().manual_seed(4)
    V = 8
    x = torch.randn(3, 2, V, generator=g)
    tg = torch.tensor([[3, 0], [3, 0]]); tl = torch.tensor([2, 2]); il = torch.tensor([2, 3])
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev()).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    assert torch.isinf(loss[0]) and loss[0] > 0
    assert abs(float(loss[1]) / ol[1] - 1) < LOSS_RTOL
    loss[1].backward()
    assert not xd.grad[:, 0].any()
    assert np.abs(xd.grad[:, 1].double().cpu().numpy() - og[:, 1]).max() < GRAD_ATOL


def test_double_backward_raises(hb):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(20, 2, 8, generator=g).to(dev()).requires_grad_(True)
    tg = torch.randint(1, 8, (2, 4), generator=g).to(dev())
    il = torch.tensor([20, 20], device=dev()); tl = torch.tensor([4, 4], device=dev())
    loss = hb.ctc_forward_score3(x, tg, il, tl, from_logits=True).sum()
    (gx,) = torch.autograd.grad(loss, x, create_graph=True)
    with pytest.raises((NotImplementedError, RuntimeError)):
        gx.pow(2).sum().backward()


def test_rnnt_joint_free_with_disagreeing_peaky_factors(hb):
    """f and g that disagree strongly about the likeliest class: the joint-free path must agree with the materialised
    joint (the reference call site) while the products exp(f - max f) exp(g - max g) are still inside the fp32 range,
    and must never return NaN (Limits)."""
    g = torch.Generator().manual_seed(12)
    N, T, U, V = 2, 12, 5, 16
    f = torch.randn(N, T, V, generator=g); gg = torch.randn(N, U + 1, V, generator=g)
    f[:, :, 3] += 30.0; gg[:, :, 9] += 30.0              # 60 nats of disagreement: representable
    tg = torch.randint(1, V, (N, U), generator=g)
    il = torch.full((N,), T); tl = torch.full((N,), U)
    fd, gd = f.to(dev()).requires_grad_(True), gg.to(dev()).requires_grad_(True)
    l1 = hb.transducer_forward_score_fg(fd, gd, tg.to(dev()), il.to(dev()), tl.to(dev()))
    joint = (f[:, :, None, :] + gg[:, None, :, :]).to(dev())
    l2 = hb.transducer_forward_score(joint, tg.to(dev()), il.to(dev()), tl.to(dev()), from_logits=True)
    assert torch.isfinite(l1).all()
    assert (l1 / l2 - 1).abs().max() < LOSS_RTOL
    l1.sum().backward()
    assert torch.isfinite(fd.grad).all() and torch.isfinite(gd.grad).all()


# ------------------------------------------------------------ strided RNN-T joints (no copy) ---
@pytest.mark.parametrize("layout", ["permuted", "padded", "unaligned"])
def test_rnnt_takes_joint_views_as_they_are(hb, oracle, layout):
    """The joint is read through its (n, t, u) strides and the gradient comes back with the strides of the view
    (`joint.contiguous()` was a silent copy of the whole lattice): a (N,U+1,T,V) buffer viewed as
    (N,T,U+1,V), a joint sliced out of a wider-U buffer, and a slice whose rows are not 16-byte aligned (plain loads
    instead of bulk copies)."""
    g = torch.Generator().manual_seed(77)
    N, T, U, V = 3, 21, 7, 16
    x = torch.randn(N, T, U + 1, V, generator=g)
    tg = torch.randint(0, V, (N, U), generator=g)
    il = torch.tensor([21, 13, 20]); tl = torch.tensor([7, 7, 3])
    ol, og = oracle.rnnt(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    if layout == "permuted":
        base = x.permute(0, 2, 1, 3).contiguous().to(dev())
        view = base.permute(0, 2, 1, 3)
    elif layout == "padded":
        base = torch.zeros(N, T, U + 4, V, device=dev()); base[:, :, :U + 1] = x.to(dev())
        view = base[:, :, :U + 1]
    else:
        base = torch.zeros(N, T, U + 1, V + 2, device=dev()); base[..., 1:V + 1] = x.to(dev())
        view = base[..., 1:V + 1]
    assert not view.is_contiguous()
    from haloop_b200 import ops
    loss, ws = ops.rnnt_fwd(view, tg.to(dev()), il.to(dev()), tl.to(dev()), True)
    gr = ops.rnnt_bwd(view, ws, torch.ones(N, device=dev()), True)
    np.testing.assert_allclose(loss.double().cpu().numpy(), ol, rtol=LOSS_RTOL)
    assert np.abs(gr.double().cpu().numpy() - og).max() < GRAD_ATOL
    if layout == "permuted":
        assert gr.stride() == view.stride(), "the gradient of a dense permuted view keeps its strides"

"""Generate tests/golden/*.npz by running the UNMODIFIED reference in float64.

Run in the build container only (it needs /root/reference):
    python oracle/make_golden.py
The reference functions (ha/ctc.py:110-174, ha/star.py:65-163,
ha/transducer.py:175-205) are dtype-generic, so the gold is the reference
itself evaluated on float64 copies of float32-representable inputs; the same
code in float32 gives the reference's own noise floor, stored as *_ref32_dev.
Gradients are w.r.t. the LOGITS:  x.requires_grad_(); f(x.log_softmax(-1)).sum().backward()
and, for the small cases, also w.r.t. the log-probs (autograd at the reference's
`emissions`/`joint` argument).  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from ha.ctc import ctc_forward_score3, ctc_reduce_mean  # noqa: E402
from ha.star import star_ctc_forward_score  # noqa: E402
from ha.transducer import transducer_forward_score  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def var_lengths(g, full, n, lo_frac=0.5):
    """SURVEY §8(d) recipe: first utterance full length, others U{full/2..full}."""
    lens = torch.randint(int(full * lo_frac), full + 1, (n,), generator=g)
    lens[0] = full
    return lens


def run_ref(fn, x32, *args, want_lp_grad=False, **kw):
    out = {}
    for name, dt in (("64", torch.float64), ("32", torch.float32)):
        x = x32.to(dt).clone().requires_grad_(True)
        lp = x.log_softmax(-1)
        if want_lp_grad:
            lp.retain_grad()
        losses = fn(lp, *args, **kw)
        losses.sum().backward()
        out["loss" + name] = losses.detach()
        out["grad" + name] = x.grad.detach()
        if want_lp_grad:
            out["lpgrad" + name] = lp.grad.detach()
    return out


def pack(case, r, x32, extra, rows=None, small=True):
    d = dict(extra)
    d["loss"] = r["loss64"].numpy()
    d["ref32_loss_dev"] = float(((r["loss32"].double() - r["loss64"]) / r["loss64"]).abs().max())
    d["ref32_grad_dev"] = float((r["grad32"].double() - r["grad64"]).abs().max())
    d["grad_abs_sum"] = float(r["grad64"].abs().sum())
    if small:
        d["x"] = x32.numpy()
        d["grad"] = r["grad64"].numpy()
        if "lpgrad64" in r:
            d["lpgrad"] = r["lpgrad64"].numpy()
    else:
        d["x_checksum"] = float(x32.double().sum())
        d["grad_rows"] = np.asarray(rows)
    return d


def ctc_like_case(kind, name, T, N, V, S, seed, var=False, repeats=False, drop=0.0, small=True,
                  star_penalty=-0.5, x_scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x32 = torch.randn(T, N, V, generator=g, dtype=torch.float32) * x_scale
    hi = 3 if repeats else V
    tg = torch.randint(1, hi, (N, S), generator=g)
    il = var_lengths(g, T, N) if var else torch.full((N,), T)
    tl = var_lengths(g, S, N) if var else torch.full((N,), S)
    if drop > 0:  # partial labels, mirrors WordDrop ha/data.py:154-169
        keep = torch.rand(N, S, generator=g) >= drop
        new = torch.zeros_like(tg)
        for n in range(N):
            k = tg[n, :tl[n]][keep[n, :tl[n]]]
            if len(k) == 0:
                k = tg[n, :1]
            new[n, :len(k)] = k
            tl[n] = len(k)
        tg = new
    if kind == "ctc":
        r = run_ref(ctc_forward_score3, x32, tg, il, tl, want_lp_grad=small)
        extra = {"reduce_mean": float(ctc_reduce_mean(r["loss64"], tl))}
    else:
        r = run_ref(star_ctc_forward_score, x32, tg, il, tl, want_lp_grad=small,
                    star_penalty=star_penalty)
        extra = {"star_penalty": star_penalty, "reduce_mean": float(ctc_reduce_mean(r["loss64"], tl))}
    extra.update(kind=kind, seed=seed, shape=np.array([T, N, V, S]), targets=tg.numpy(),
                 in_len=il.numpy(), tgt_len=tl.numpy(), x_scale=x_scale)
    rows = [0, N - 1]
    d = pack(name, r, x32, extra, rows=rows, small=small)
    if not small:
        d["grad_sub"] = r["grad64"][:, rows, :].to(torch.float32).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: loss[0]={float(r['loss64'][0]):.8f} ref32 dev loss {d['ref32_loss_dev']:.2e} "
          f"grad {d['ref32_grad_dev']:.2e}")


def rnnt_case(name, N, T, U, V, seed, var=False, zero_targets=False, small=True):
    assert 2 ** round(np.log2(T)) >= T, "reference scan width bug (SURVEY finding 4)"
    g = torch.Generator().manual_seed(seed)
    x32 = torch.randn(N, T, U + 1, V, generator=g, dtype=torch.float32)
    tg = torch.randint(0 if zero_targets else 1, V, (N, U), generator=g)
    il = var_lengths(g, T, N) if var else torch.full((N,), T)
    tl = var_lengths(g, U, N) if var else torch.full((N,), U)
    r = run_ref(transducer_forward_score, x32, tg, il, tl, want_lp_grad=small)
    extra = dict(kind="rnnt", seed=seed, shape=np.array([N, T, U, V]), targets=tg.numpy(),
                 in_len=il.numpy(), tgt_len=tl.numpy())
    rows = [0, N - 1]
    d = pack(name, r, x32, extra, rows=rows, small=small)
    if not small:
        d["grad_sub"] = r["grad64"][rows].to(torch.float32).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: loss[0]={float(r['loss64'][0]):.8f} ref32 dev loss {d['ref32_loss_dev']:.2e} "
          f"grad {d['ref32_grad_dev']:.2e}")


def kat_appendix_d():
    """SURVEY.md Appendix D seeded known answers, regenerated and asserted."""
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(12, 3, 6, generator=g, dtype=torch.float64)
    tg = torch.randint(1, 6, (3, 4), generator=g)
    il, tl = torch.tensor([12, 9, 7]), torch.tensor([4, 3, 2])
    out = {"x": x.numpy(), "targets": tg.numpy(), "in_len": il.numpy(), "tgt_len": tl.numpy()}
    for nm, fn, kw in (("ctc", ctc_forward_score3, {}),
                       ("star", star_ctc_forward_score, {"star_penalty": -0.5})):
        xx = x.clone().requires_grad_(True)
        l = fn(xx.log_softmax(-1), tg, il, tl, **kw)
        l.sum().backward()
        out[nm + "_loss"] = l.detach().numpy(); out[nm + "_grad"] = xx.grad.numpy()
    assert np.allclose(out["ctc_loss"], [12.07453682, 11.34123406, 8.05409432], atol=1e-7)
    assert np.allclose(out["star_loss"], [8.56281179, 6.97394701, 6.39507674], atol=1e-7)
    g = torch.Generator().manual_seed(4321)
    j = torch.randn(3, 8, 5, 6, generator=g, dtype=torch.float64)
    tg2 = torch.randint(1, 6, (3, 4), generator=g)
    jl, tl2 = torch.tensor([8, 6, 5]), torch.tensor([4, 2, 3])
    jj = j.clone().requires_grad_(True)
    l = transducer_forward_score(jj.log_softmax(-1), tg2, jl, tl2)
    l.sum().backward()
    assert np.allclose(l.detach().numpy(), [15.62174349, 10.61215748, 10.60625096], atol=1e-7)
    out.update(joint=j.numpy(), rnnt_targets=tg2.numpy(), rnnt_in_len=jl.numpy(),
               rnnt_tgt_len=tl2.numpy(), rnnt_loss=l.detach().numpy(), rnnt_grad=jj.grad.numpy())
    np.savez_compressed(os.path.join(OUT, "kat_appendix_d.npz"), **out)
    print("kat_appendix_d ok")


def rnnt_fg_case(name, N, T, U, V, seed, var=False, zero_targets=False):
    """The reference's own additive joint (ha/recognizer.py:114) differentiated w.r.t. its two factors."""
    assert 2 ** round(np.log2(T)) >= T, "reference scan width bug (SURVEY finding 4)"
    g_ = torch.Generator().manual_seed(seed)
    f32 = torch.randn(N, T, V, generator=g_, dtype=torch.float32)
    g32 = torch.randn(N, U + 1, V, generator=g_, dtype=torch.float32)
    tg = torch.randint(0 if zero_targets else 1, V, (N, U), generator=g_)
    il = var_lengths(g_, T, N) if var else torch.full((N,), T)
    tl = var_lengths(g_, U, N) if var else torch.full((N,), U)
    f = f32.double().requires_grad_(True); g = g32.double().requires_grad_(True)
    joint = f[:, :, None, :] + g[:, None, :, :]
    losses = transducer_forward_score(joint.log_softmax(-1), tg, il, tl)
    losses.sum().backward()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="rnnt_fg", seed=seed, f=f32.numpy(), g=g32.numpy(),
                        targets=tg.numpy(), in_len=il.numpy(), tgt_len=tl.numpy(), loss=losses.detach().numpy(),
                        grad_f=f.grad.numpy(), grad_g=g.grad.numpy())
    print(f"{name}: loss[0]={float(losses[0]):.8f}")


def main_fg():
    os.makedirs(OUT, exist_ok=True)
    rnnt_fg_case("rnntfg_small_var", 5, 16, 6, 10, seed=51, var=True)
    rnnt_fg_case("rnntfg_zero_labels", 4, 8, 5, 6, seed=52, var=True, zero_targets=True)
    rnnt_fg_case("rnntfg_medium", 3, 64, 20, 40, seed=53, var=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    kat_appendix_d()
    # CTC (ha/ctc.py)
    ctc_like_case("ctc", "ctc_small", 30, 4, 12, 6, seed=11)
    ctc_like_case("ctc", "ctc_small_var", 50, 6, 20, 10, seed=12, var=True)
    ctc_like_case("ctc", "ctc_repeats", 40, 5, 5, 12, seed=13, var=True, repeats=True)
    ctc_like_case("ctc", "ctc_peaky", 60, 4, 24, 14, seed=14, var=True, x_scale=4.0)
    ctc_like_case("ctc", "ctc_slots", 160, 3, 40, 70, seed=15, var=True)   # > 32 label pairs
    ctc_like_case("ctc", "ctc_c1", 200, 8, 256, 50, seed=0, small=False)
    ctc_like_case("ctc", "ctc_c1_var", 200, 8, 256, 50, seed=1, var=True, small=False)
    # star-CTC (ha/star.py)
    ctc_like_case("star", "star_small", 30, 4, 12, 6, seed=21)
    ctc_like_case("star", "star_small_var", 50, 6, 20, 10, seed=22, var=True)
    ctc_like_case("star", "star_repeats", 40, 5, 5, 12, seed=23, var=True, repeats=True)
    ctc_like_case("star", "star_pen0", 40, 4, 16, 8, seed=24, var=True, star_penalty=0.0)
    ctc_like_case("star", "star_drop", 80, 6, 24, 16, seed=25, var=True, drop=0.4)
    ctc_like_case("star", "star_slots", 160, 3, 40, 70, seed=26, var=True)
    ctc_like_case("star", "star_c1", 200, 8, 256, 50, seed=2, small=False)
    ctc_like_case("star", "star_c1_drop", 200, 8, 256, 50, seed=3, var=True, drop=0.4, small=False)
    # RNN-T (ha/transducer.py)
    rnnt_case("rnnt_test_batched_shape", 13, 7, 4, 6, seed=42, zero_targets=True)
    rnnt_case("rnnt_small_var", 5, 16, 6, 10, seed=31, var=True)
    rnnt_case("rnnt_medium", 4, 64, 12, 32, seed=32, var=True)
    rnnt_case("rnnt_wide", 2, 128, 40, 48, seed=33, var=True, small=False)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fg":
        main_fg()             # only the joint-free RNN-T cases (added after the first set was committed)
    else:
        main()
        main_fg()

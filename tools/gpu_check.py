"""Diagnostic run for a fresh GPU box: exercises every entry point at small sizes and prints the
errors against the CPU oracle without stopping at the first failure.  Not a test; see tests/."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import haloop_b200 as hb  # noqa: E402
from oracle import oracle  # noqa: E402

dev = torch.device("cuda:0")


def report(name, loss, ol, grad, og):
    lo = loss.detach().double().cpu().numpy()
    fin = np.isfinite(ol)
    dl = np.abs(lo[fin] / ol[fin] - 1).max() if fin.any() else 0.0
    dg = np.abs(grad.double().cpu().numpy() - og).max()
    flag = "OK " if (dl < 1e-4 and dg < 1e-5) else "BAD"
    print(f"[{flag}] {name}: loss rel {dl:.2e} grad abs {dg:.2e}  loss[:3]={lo[:3]} oracle={ol[:3]}", flush=True)


def run(name, fn):
    try:
        t = time.time()
        fn()
        torch.cuda.synchronize()
        print(f"      ({name} took {time.time() - t:.2f}s)", flush=True)
    except Exception:
        print(f"[EXC] {name}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("CUDA context is dead:", e, flush=True)
            sys.exit(3)


def ctc_case(T, N, V, S, seed=0, var=True, from_logits=True, star=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.randint(T // 2, T + 1, (N,), generator=g) if var else torch.full((N,), T)
    tl = torch.randint(S // 2, S + 1, (N,), generator=g) if var else torch.full((N,), S)
    il[0] = T; tl[0] = S
    for n in range(N):
        tg[n, tl[n]:] = 0
    xin = x if from_logits else x.double().log_softmax(-1).float()
    xd = xin.to(dev).requires_grad_(True)
    if star:
        loss = hb.star_ctc_forward_score(xd, tg.to(dev), il.to(dev), tl.to(dev), star_penalty=-0.5, from_logits=from_logits)
        ol, og = oracle.star(xin.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5, from_logits=from_logits)
    else:
        loss = hb.ctc_forward_score3(xd, tg.to(dev), il.to(dev), tl.to(dev), from_logits=from_logits)
        ol, og = oracle.ctc(xin.numpy(), tg.numpy(), il.numpy(), tl.numpy(), from_logits=from_logits)
    loss.sum().backward()
    report(f"{'star' if star else 'ctc'} T={T} N={N} V={V} S={S} logits={from_logits}", loss, ol, xd.grad, og)


def rnnt_case(N, T, U, V, seed=0, from_logits=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, T, U + 1, V, generator=g)
    tg = torch.randint(1, V, (N, U), generator=g)
    il = torch.randint(T // 2, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(U // 2, U + 1, (N,), generator=g); tl[0] = U
    xin = x if from_logits else x.double().log_softmax(-1).float()
    xd = xin.to(dev).requires_grad_(True)
    loss = hb.transducer_forward_score(xd, tg.to(dev), il.to(dev), tl.to(dev), from_logits=from_logits)
    loss.sum().backward()
    ol, og = oracle.rnnt(xin.numpy(), tg.numpy(), il.numpy(), tl.numpy(), from_logits=from_logits)
    report(f"rnnt N={N} T={T} U={U} V={V} logits={from_logits}", loss, ol, xd.grad, og)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ["ctc", "star", "rnnt"]
    if "ctc" in which:
        run("ctc tiny", lambda: ctc_case(12, 3, 8, 4))
        run("ctc small", lambda: ctc_case(64, 4, 32, 9))
        run("ctc lp", lambda: ctc_case(64, 4, 32, 9, from_logits=False))
        run("ctc odd V", lambda: ctc_case(50, 3, 37, 11))
        run("ctc slots", lambda: ctc_case(300, 4, 64, 100))
        run("ctc C1", lambda: ctc_case(200, 8, 256, 50, var=False))
        run("ctc long", lambda: ctc_case(1500, 4, 1024, 300, var=False))
    if "star" in which:
        run("star tiny", lambda: ctc_case(12, 3, 8, 4, star=True))
        run("star small", lambda: ctc_case(64, 4, 32, 9, star=True))
        run("star lp", lambda: ctc_case(64, 4, 32, 9, from_logits=False, star=True))
        run("star odd V", lambda: ctc_case(50, 3, 37, 11, star=True))
        run("star slots", lambda: ctc_case(300, 4, 64, 100, star=True))
        run("star C3-ish", lambda: ctc_case(1000, 4, 512, 200, var=False, star=True))
    if "rnnt" in which:
        run("rnnt tiny", lambda: rnnt_case(2, 6, 3, 8))
        run("rnnt small", lambda: rnnt_case(3, 20, 6, 16))
        run("rnnt lp", lambda: rnnt_case(3, 20, 6, 16, from_logits=False))
        run("rnnt odd V", lambda: rnnt_case(3, 21, 7, 37))
        run("rnnt wide", lambda: rnnt_case(2, 40, 70, 32))
        run("rnnt C4-ish", lambda: rnnt_case(2, 200, 100, 256))

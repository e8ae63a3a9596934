"""GPU check of the fused CTC path against the float64 oracle over the shapes that exercise every
code path (trellis warp counts, phantom groups, T = 1, L = 0, repeats, label 0, log-prob mode,
grad_output != 1, permuted views), then a quick timing of BASELINE config 2.

  python tools/ctc2_check.py [--time]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb  # noqa: E402
from oracle import oracle  # noqa: E402

dev = torch.device("cuda:0")


def case(T, N, V, S, seed=0, scale=1.0, var=True, from_logits=True, gout=False, permuted=False, special=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g) * scale
    tg = torch.randint(1, V, (N, max(S, 1)), generator=g)
    if S == 0:
        tg = tg[:, :0]
    il = torch.full((N,), T, dtype=torch.int64)
    tl = torch.full((N,), S, dtype=torch.int64)
    if var and N > 1:
        il[1:] = torch.randint(max(T // 2, 1), T + 1, (N - 1,), generator=g)
        tl[1:] = torch.randint(S // 2, S + 1, (N - 1,), generator=g)
    if special and S >= 6:
        tg[:, 2] = tg[:, 1]
        tg[0, 4] = 0
    # keep every utterance feasible
    for n in range(N):
        need = int(tl[n]) + sum(1 for k in range(1, int(tl[n])) if tg[n, k] == tg[n, k - 1] or tg[n, k] == 0)
        if il[n] < need:
            il[n] = min(T, need)
            if il[n] < need:
                tl[n] = max(0, int(tl[n]) - (need - T))
    if not from_logits:
        x = x.log_softmax(-1)
    go = (torch.rand(N, generator=g) + 0.5) if gout else None
    if permuted:
        xd = x.permute(1, 0, 2).contiguous().to(dev).permute(1, 0, 2).requires_grad_(True)
    else:
        xd = x.to(dev).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev), il.to(dev), tl.to(dev), from_logits=from_logits)
    if go is None:
        loss.sum().backward()
    else:
        (loss * go.to(dev)).sum().backward()
    torch.cuda.synchronize()
    ol, og = oracle.ctc(x.numpy(), tg.numpy() if S else np.zeros((N, 1), np.int64), il.numpy(), tl.numpy(),
                        from_logits=from_logits, grad_out=None if go is None else go.numpy())
    l = loss.detach().cpu().double().numpy()
    fin = np.isfinite(ol)
    dl = float(np.abs(l[fin] / ol[fin] - 1).max()) if fin.any() else 0.0
    dg = float(np.abs(xd.grad.cpu().double().numpy() - og).max())
    same_inf = bool(np.array_equal(np.isfinite(l), fin))
    return dl, dg, same_inf


def main():
    cases = [
        dict(T=64, N=4, V=32, S=9),
        dict(T=1, N=2, V=8, S=1, var=False, special=False),
        dict(T=2, N=2, V=8, S=1, var=False, special=False),
        dict(T=40, N=3, V=16, S=0),
        dict(T=33, N=4, V=12, S=16),
        dict(T=200, N=8, V=256, S=50),
        dict(T=200, N=8, V=256, S=50, from_logits=False),
        dict(T=200, N=8, V=256, S=50, gout=True, permuted=True),
        dict(T=300, N=6, V=64, S=120),         # NL = 32: exactly one warp
        dict(T=300, N=6, V=64, S=121),         # NL = 33: two warps
        dict(T=500, N=5, V=128, S=200),
        dict(T=700, N=4, V=512, S=300),
        dict(T=1100, N=3, V=64, S=500),
        dict(T=2100, N=2, V=32, S=1000),
        dict(T=400, N=4, V=64, S=60, scale=5.0),
        dict(T=400, N=4, V=64, S=60, scale=6.0, from_logits=False),
        dict(T=1500, N=6, V=1024, S=300),
        dict(T=16000, N=2, V=16, S=4, var=False, special=False),      # T >> L: adjacent states 30+ nats apart
        dict(T=30000, N=1, V=16, S=40, var=False),
        dict(T=1500, N=6, V=1024, S=300, scale=3.0),
    ]
    bad = 0
    for c in cases:
        try:
            dl, dg, si = case(**c)
            ok = dl < 1e-4 and dg < 1e-5 and si
            bad += not ok
            print(("ok  " if ok else "FAIL"), c, f"loss rel {dl:.2e} grad abs {dg:.2e} inf-match {si}", flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("EXC ", c, repr(e)[:300], flush=True)
            if "CUDA" in repr(e) or "cuda" in repr(e):
                break
    print("failures:", bad)
    if "--time" in sys.argv and bad == 0:
        T, N, V, S = 1500, 256, 1024, 300
        g = torch.Generator().manual_seed(0)
        x = torch.randn(T, N, V, generator=g).to(dev).requires_grad_(True)
        tg = torch.randint(1, V, (N, S), generator=g).to(dev)
        il = torch.full((N,), T, dtype=torch.int64, device=dev)
        tl = torch.full((N,), S, dtype=torch.int64, device=dev)
        from haloop_b200 import ops
        go = torch.ones(N, device=dev)
        for _ in range(3):
            loss, ws = ops.ctc_fwd(x.detach(), tg, il, tl, True)
            gx = ops.ctc_bwd(x.detach(), ws, go, S, True)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tb = 0.0
        K = 20
        for _ in range(K):
            e[0].record()
            loss, ws = ops.ctc_fwd(x.detach(), tg, il, tl, True)
            e[1].record()
            gx = ops.ctc_bwd(x.detach(), ws, go, S, True)
            e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        print(f"C2 fwd {tf / K:.3f} ms  bwd {tb / K:.3f} ms  total {(tf + tb) / K:.3f} ms  "
              f"roofline {8 * V * T * N / ((tf + tb) / K * 1e-3) / 6457.4e9:.3f}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

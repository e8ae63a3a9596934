"""Locate star-CTC gradient errors against the oracle (diagnostic, not a test)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb
from oracle import oracle
dev = torch.device("cuda:0")
np.set_printoptions(precision=5, suppress=True, linewidth=200)

def case(T, N, V, S, seed=0, var=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.randint(T // 2, T + 1, (N,), generator=g) if var else torch.full((N,), T)
    tl = torch.randint(S // 2, S + 1, (N,), generator=g) if var else torch.full((N,), S)
    il[0] = T; tl[0] = S
    for n in range(N):
        tg[n, tl[n]:] = 0
    xd = x.to(dev).requires_grad_(True)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev), il.to(dev), tl.to(dev), star_penalty=-0.5, from_logits=True)
    loss.sum().backward()
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5)
    err = np.abs(xd.grad.cpu().numpy() - og)
    print(f"T={T} N={N} V={V} S={S}: max err {err.max():.3e}; il={il.tolist()} tl={tl.tolist()}")
    print(" per-utterance max err:", err.max(axis=(0, 2)))
    print(" per-class max err:", err.max(axis=(0, 1))[:16])
    t, n, c = np.unravel_index(err.argmax(), err.shape)
    print(f" worst at t={t} n={n} c={c}; targets[n]={tg[n].tolist()}")
    for tt in sorted(set([0, 1, max(t - 1, 0), t, min(t + 1, T - 1), int(il[n]) - 1])):
        print(f"  t={tt} gpu   ", xd.grad[tt, n].cpu().numpy()[:12])
        print(f"  t={tt} oracle", og[tt, n][:12])
    print(" per-t max err for worst n:", err[:, n].max(axis=1)[:40])

case(12, 3, 8, 4)
case(64, 4, 32, 9)
case(12, 1, 8, 4, var=False)
case(12, 1, 8, 3, var=False)
case(13, 2, 12, 5, var=False)

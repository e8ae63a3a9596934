"""Star-CTC: compare a batched run against per-utterance solo runs (same inputs) to expose
cross-utterance interference; print G (occ[2]) per frame from the workspace."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb
from haloop_b200 import ops
from oracle import oracle
dev = torch.device("cuda:0")
np.set_printoptions(precision=5, suppress=True, linewidth=220)

def run(x, tg, il, tl):
    xd = x.to(dev).requires_grad_(True)
    loss = hb.star_ctc_forward_score(xd, tg.to(dev), il.to(dev), tl.to(dev), star_penalty=-0.5, from_logits=True)
    loss.sum().backward()
    return loss.detach().cpu().numpy(), xd.grad.cpu().numpy()

def case(T, N, V, S, seed=0, var=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g)
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.randint(T // 2, T + 1, (N,), generator=g) if var else torch.full((N,), T)
    tl = torch.randint(S // 2, S + 1, (N,), generator=g) if var else torch.full((N,), S)
    il[0] = T; tl[0] = S
    for n in range(N):
        tg[n, tl[n]:] = 0
    ol, og = oracle.star(x.numpy(), tg.numpy(), il.numpy(), tl.numpy(), star_penalty=-0.5)
    for rep in range(3):
        l, gr = run(x, tg, il, tl)
        err = np.abs(gr - og).max(axis=(0, 2))
        print(f"T={T} N={N} V={V} S={S} batched rep{rep}: per-utt err {err}")
    for n in range(N):
        l1, g1 = run(x[:, n:n + 1].contiguous(), tg[n:n + 1], il[n:n + 1], tl[n:n + 1])
        e = np.abs(g1[:, 0] - og[:, n]).max()
        print(f"   solo n={n}: err {e:.3e}")
    # pairs
    if N >= 2:
        for a in range(N):
            for b in range(N):
                if a == b: continue
                idx = [a, b]
                l2, g2 = run(x[:, idx].contiguous(), tg[idx], il[idx], tl[idx])
                e = [np.abs(g2[:, i] - og[:, idx[i]]).max() for i in range(2)]
                print(f"   pair {idx}: err {e[0]:.3e} {e[1]:.3e}")

case(12, 3, 8, 4)
case(13, 2, 12, 5, var=False)

import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb
from oracle import oracle
dev = torch.device("cuda:0")
np.set_printoptions(precision=6, suppress=True, linewidth=220)
def case(seed, T, N, V, S, scale):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, N, V, generator=g) * scale
    tg = torch.randint(1, V, (N, S), generator=g)
    il = torch.randint(T // 2, T + 1, (N,), generator=g); il[0] = T
    tl = torch.randint(S // 2, S + 1, (N,), generator=g); tl[0] = S
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev), il.to(dev), tl.to(dev), from_logits=True)
    loss.sum().backward()
    gr = xd.grad.cpu().numpy()
    err = np.abs(gr - og)
    print(f"scale={scale}: loss gpu {loss.detach().cpu().numpy()} oracle {ol}")
    print(" max err", err.max(), "per utt", err.max(axis=(0, 2)))
    t, n, c = np.unravel_index(err.argmax(), err.shape)
    print(f" worst t={t} n={n} c={c} (il={il[n]}, tl={tl[n]}) gpu {gr[t,n,c]:.7f} oracle {og[t,n,c]:.7f}; target classes {sorted(set(tg[n,:tl[n]].tolist()))[:12]}")
    print(" row gpu   ", gr[t, n, :12]); print(" row oracle", og[t, n, :12])
    pt = err[:, n].max(axis=1); idx = np.argsort(-pt)[:10]
    print(" worst frames", sorted(idx.tolist()), pt[sorted(idx.tolist())])
    # is the error in the softmax part or the occupancy part?  non-target classes carry softmax only
    nt = [c for c in range(V) if c not in set(tg[n, :tl[n]].tolist()) and c != 0]
    print(" max err on non-target classes (softmax only):", err[:, n][:, nt].max() if nt else None)
for sc in (1.0, 3.0, 5.0, 8.0):
    case(500, 400, 4, 32, 40, sc)

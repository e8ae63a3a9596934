import sys, numpy as np, torch
sys.path.insert(0, '.')
import haloop_b200 as hb
from oracle import oracle
dev = 'cuda'
N, T, U, V = 2, 130, 70, 33
gen = torch.Generator().manual_seed(500 + T)
f = torch.randn(N, T, V, generator=gen); g = torch.randn(N, U + 1, V, generator=gen)
tg = torch.randint(1, V, (N, U), generator=gen)
il = torch.randint(T // 2, T + 1, (N,), generator=gen); il[0] = T
tl = torch.randint(U // 2, U + 1, (N,), generator=gen); tl[0] = U
print(il, tl)
ol, ogf, ogg = oracle.rnnt_fg(f.numpy(), g.numpy(), tg.numpy(), il.numpy(), tl.numpy())
fd = f.to(dev).requires_grad_(True); gd = g.to(dev).requires_grad_(True)
loss = hb.transducer_forward_score_fg(fd, gd, tg.to(dev), il.to(dev), tl.to(dev))
loss.sum().backward()
print('loss', loss.detach().cpu().numpy(), ol)
ef = np.abs(fd.grad.double().cpu().numpy() - ogf); eg = np.abs(gd.grad.double().cpu().numpy() - ogg)
print('ef max', ef.max(), 'eg max', eg.max())
for n in range(N):
    print('n', n, 'ef rows>1e-5:', np.where(ef[n].max(1) > 1e-5)[0][:20], 'cols', np.where(ef[n].max(0) > 1e-5)[0][:20])
    print('n', n, 'eg rows>1e-5:', np.where(eg[n].max(1) > 1e-5)[0][:20], 'cols', np.where(eg[n].max(0) > 1e-5)[0][:20])
    print(' ef max per n', ef[n].max(), eg[n].max())

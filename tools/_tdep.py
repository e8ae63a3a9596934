import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import numpy as np, torch
import haloop_b200 as hb
from oracle import oracle
dev = torch.device('cuda:0')
for (T, V, S) in [(3000,16,40),(8000,16,40),(16000,16,40),(20000,16,40),(30000,16,40),(30000,64,40),(30000,16,4)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(T, 1, V, generator=g); tg = torch.randint(1, V, (1, S), generator=g)
    il = torch.tensor([T]); tl = torch.tensor([S])
    ol, og = oracle.ctc(x.numpy(), tg.numpy(), il.numpy(), tl.numpy())
    xd = x.to(dev).requires_grad_(True)
    loss = hb.ctc_forward_score3(xd, tg.to(dev), il.to(dev), tl.to(dev), from_logits=True)
    loss.sum().backward()
    gerr = np.abs(xd.grad.cpu().double().numpy() - og)
    print(T, V, S, float(loss), ol[0], abs(float(loss)/ol[0]-1), gerr.max(), int(gerr.reshape(T,-1).max(1).argmax()))

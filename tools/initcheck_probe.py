"""compute-sanitizer --tool initcheck probe: the same small CTC loss + gradient once with a
class count the TMA bulk-copy path takes (V = 32) and once with one it cannot take (V = 37, plain stores), the
gradient copied to the host each time.  usage: compute-sanitizer --tool initcheck python tools/initcheck_probe.py 32|37"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb
V = int(sys.argv[1])
g = torch.Generator().manual_seed(0)
T, N, S = 40, 3, 9
x = torch.randn(T, N, V, generator=g).cuda().requires_grad_(True)
tg = torch.randint(1, V, (N, S), generator=g).cuda()
il = torch.tensor([T, T - 7, T // 2]).cuda(); tl = torch.tensor([S, S // 2, 1]).cuda()
hb.ctc_forward_score3(x, tg, il, tl, from_logits=True).sum().backward()
print("V", V, "grad checksum", float(x.grad.cpu().double().abs().sum()))

"""A few small loss+grad calls of every kind, for compute-sanitizer runs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import haloop_b200 as hb
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "star":          # the fused star-CTC kernels only: 1 .. 4 trellis warps, repeats, L = 0 / S, T_n < T
    for (T, N, V, S) in [(40, 3, 32, 9), (150, 2, 24, 70), (260, 2, 8, 200), (700, 1, 12, 400)]:
        x = torch.randn(T, N, V, generator=g).to(dev).requires_grad_(True)
        tg = torch.randint(1, min(V, 6), (N, S), generator=g).to(dev)
        il = torch.tensor([T, T - 7][:N] + [T // 2] * max(0, N - 2)).to(dev); tl = torch.tensor([S, S // 2][:N] + [0] * max(0, N - 2)).to(dev)
        hb.star_ctc_forward_score(x, tg, il, tl, from_logits=True).sum().backward()
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)
for (T, N, V, S) in [(40, 3, 32, 9), (150, 2, 24, 70), (33, 2, 37, 5)]:
    x = torch.randn(T, N, V, generator=g).to(dev).requires_grad_(True)
    tg = torch.randint(1, V, (N, S), generator=g).to(dev)
    il = torch.tensor([T, T - 7][:N] + [T // 2] * max(0, N - 2)).to(dev); tl = torch.tensor([S, S // 2][:N] + [1] * max(0, N - 2)).to(dev)
    hb.ctc_forward_score3(x, tg, il, tl, from_logits=True).sum().backward()
    x.grad = None
    hb.star_ctc_forward_score(x, tg, il, tl, from_logits=True).sum().backward()
for (N, T, U, V) in [(2, 17, 6, 16), (2, 12, 40, 37)]:
    j = torch.randn(N, T, U + 1, V, generator=g).to(dev).requires_grad_(True)
    tg = torch.randint(1, V, (N, U), generator=g).to(dev)
    il = torch.tensor([T, T - 3]).to(dev); tl = torch.tensor([U, U // 2]).to(dev)
    hb.transducer_forward_score(j, tg, il, tl, from_logits=True).sum().backward()
    f = torch.randn(N, T + 60, V, generator=g).to(dev).requires_grad_(True)          # several GEMM tiles
    gg = torch.randn(N, U + 1, V, generator=g).to(dev).requires_grad_(True)
    hb.transducer_forward_score_fg(f, gg, tg, torch.tensor([T + 60, T]).to(dev), tl).sum().backward()
j = torch.randn(1, 4, 200, 8, generator=g).to(dev).requires_grad_(True)                # the 1024-thread lattice
hb.transducer_forward_score(j, torch.randint(0, 8, (1, 199), generator=g).to(dev), torch.tensor([4]).to(dev),
                            torch.tensor([199]).to(dev), from_logits=True).sum().backward()
lp = torch.randn(3, 50, 20, generator=g).log_softmax(-1).to(dev)
hb.greedy_decode(lp, torch.tensor([50, 40, 3]).to(dev))
hb.ctc_viterbi_align(lp.permute(1, 0, 2), torch.randint(1, 20, (3, 8), generator=g).to(dev), torch.tensor([50, 40, 30]).to(dev), torch.tensor([8, 5, 2]).to(dev))
hb.ctc_beam_search_decode_logits(lp, 4, torch.tensor([50, 40, 3]).to(dev), graves=True)
hb.ctc_beam_search_decode_logits(lp[0, :9], 3)
# fused classifier head + CTC (tensor-core GEMM engine; V = 40 also takes the joint-free RNN-T through it above when V % 16 == 0)
for (N, T, D, V, S) in [(3, 70, 64, 40, 7), (2, 130, 100, 260, 11)]:
    h = torch.randn(N, T, D, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(V, D, generator=g) / D ** 0.5).to(dev).requires_grad_(True)
    b = torch.randn(V, generator=g).to(dev).requires_grad_(True)
    tg = torch.randint(1, V, (N, S), generator=g).to(dev)
    il = torch.tensor([T, T - 9, T // 2][:N]).to(dev); tl = torch.tensor([S, S // 2, 1][:N]).to(dev)
    hb.linear_ctc_forward_score(h, W, b, tg, il, tl).sum().backward()
for (N, T, U, V) in [(2, 140, 9, 32)]:                                                 # joint-free RNN-T on the engine (V % 16 == 0)
    f = torch.randn(N, T, V, generator=g).to(dev).requires_grad_(True)
    gg = torch.randn(N, U + 1, V, generator=g).to(dev).requires_grad_(True)
    tg = torch.randint(1, V, (N, U), generator=g).to(dev)
    hb.transducer_forward_score_fg(f, gg, tg, torch.tensor([T, T - 30]).to(dev), torch.tensor([U, U // 2]).to(dev)).sum().backward()
torch.cuda.synchronize()
print("done")

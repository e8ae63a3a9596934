"""Forward / backward time of the fused star-CTC kernels against the batch size (BASELINE config 3 shapes): 74 utterances
= one CTA per SM, 128 = config 3 (108 SMs hold two CTAs), 148 = two per SM, 222 = three.  Shows how much of a step is
contention between the CTAs of an SM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import ops  # noqa: E402

T, V, U = 1000, 512, 200
dev = torch.device("cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [37, 74, 128, 148, 222, 296]:
    g = torch.Generator().manual_seed(B)
    xs = [torch.randn(T, B, V, generator=g).to(dev) for _ in range(3)]
    tg = torch.randint(1, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, device=dev); tl = torch.full((B,), U, device=dev)
    go = torch.ones(B, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
    for it in range(25):
        x = xs[it % 3]
        if it >= 5: ev[it - 5][0].record()
        loss, ws = ops.star_fwd(x, tg, il, tl, -0.5, True)
        if it >= 5: ev[it - 5][1].record()
        gr = ops.star_bwd(x, ws, go, U, True)
        if it >= 5: ev[it - 5][2].record()
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev); b = sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev)
    print(f"B={B:4d} ({2 * B} CTAs): fwd {f:.4f} ms  bwd {b:.4f} ms  step {f + b:.4f} ms  -> {8 * V * B * T / (f + b) / 1e6 / 6457.4:.3f} of the HBM roofline")

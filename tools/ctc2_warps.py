"""How much of a fused CTC step the trellis warps cost: BASELINE config 2 shapes with target lengths either side of a
warp boundary (247 labels = 64 lanes = 2 trellis warps, 249 labels = 65 lanes = 3 warps; the row work is the same)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import ops  # noqa: E402

B, T, V = 256, 1500, 1024
dev = torch.device("cuda:0")
for U in [int(a) for a in sys.argv[1:]] or [247, 249, 300]:
    g = torch.Generator().manual_seed(U)
    xs = [torch.randn(T, B, V, generator=g).to(dev) for _ in range(2)]
    tg = torch.randint(1, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, device=dev); tl = torch.full((B,), U, device=dev)
    go = torch.ones(B, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(20)]
    for it in range(25):
        x = xs[it % 2]
        if it >= 5: ev[it - 5][0].record()
        loss, ws = ops.ctc_fwd(x, tg, il, tl, True)
        if it >= 5: ev[it - 5][1].record()
        gr = ops.ctc_bwd(x, ws, go, U, True)
        if it >= 5: ev[it - 5][2].record()
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in ev) / len(ev); b = sum(e[1].elapsed_time(e[2]) for e in ev) / len(ev)
    print(f"U={U}: fwd {f:.4f} ms  bwd {b:.4f} ms  step {f + b:.4f} ms")

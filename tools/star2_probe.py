"""Per-step cycle budget of the fused star-CTC kernels from the clock64 probes of a -DHAB_STAR2_PROBE build
(HA_B200_SO=haloop_b200/libha_b200_probe.so python tools/star2_probe.py): where a trellis warp and a row warp of one
CTA pair spend their time (BASELINE config 3 shapes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import ops  # noqa: E402

B, T, V, U = 128, 1000, 512, 200
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(T, B, V, generator=g).to(dev)
tg = torch.randint(1, V, (B, U), generator=g).to(dev)
il = torch.full((B,), T, device=dev); tl = torch.full((B,), U, device=dev)
for it in range(3):
    loss, ws = ops.star_fwd(x, tg, il, tl, -0.5, True)
    h_fwd = ws[:256].view(torch.float32).clone().cpu()
    gr = ops.star_bwd(x, ws, torch.ones(B, device=dev), U, True)
    torch.cuda.synchronize()
h = ws[:256].view(torch.float32).cpu()
names_f = ["em wait", "compute", "barrier", "loop"]
names_b = ["em wait", "stored-row wait", "compute", "barrier", "loop"]
names_r = ["row wait", "pre", "occ wait", "post", "store-read wait", "other"]
for b in (0, 1):
    print(f"CTA {b} (dir {b})")
    print("  fwd trellis, cycles/step:", {n: round(float(h_fwd[16 + 8 * b + i])) for i, n in enumerate(names_f)})
    print("  bwd trellis, cycles/step:", {n: round(float(h[32 + 8 * b + i])) for i, n in enumerate(names_b)})
    print("  bwd row warp 0, cycles/row:", {n: round(float(h[48 + 8 * b + i])) for i, n in enumerate(names_r)})

"""Time the CTC forward (prep + rows + trellis) for a few shapes to separate per-step latency from
throughput effects (diagnostic)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import ops
dev = torch.device("cuda:0")

def run(B, T, V, U, reps=5):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, V, generator=g).to(dev)
    tg = torch.randint(1, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, dtype=torch.int64, device=dev); tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    xv = x.permute(1, 0, 2)
    for _ in range(2):
        ops.ctc_fwd(xv, tg, il, tl, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.ctc_fwd(xv, tg, il, tl, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"B={B:4d} T={T:5d} V={V:5d} U={U:4d} W={os.environ.get('HA_B200_TRELLIS_W','-')}: fwd {ms:7.3f} ms  -> {ms*1e3/T:6.3f} us/step", flush=True)

for B, T, V, U in [(256, 1500, 1024, 300), (128, 1500, 1024, 300), (64, 1500, 1024, 300), (16, 1500, 1024, 300),
                   (256, 750, 1024, 300), (256, 1500, 1024, 100), (256, 1500, 1024, 30), (16, 1500, 64, 30), (16, 1500, 64, 300)]:
    run(B, T, V, U)

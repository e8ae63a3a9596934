import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import ops
dev = torch.device("cuda:0")
def run(B, T, V, U, reps=10):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, V, generator=g).to(dev)
    tg = torch.randint(1, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, dtype=torch.int64, device=dev); tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    xv = x.permute(1, 0, 2)
    for _ in range(2): ops.ctc_fwd(xv, tg, il, tl, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.ctc_fwd(xv, tg, il, tl, True)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} T={T} V={V} U={U} W={os.environ.get('HA_B200_TRELLIS_W','-')}: fwd {e0.elapsed_time(e1)/reps:.3f} ms", flush=True)
run(256, 1500, 1024, 300)
run(128, 1000, 512, 200)

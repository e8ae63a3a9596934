import os, sys, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haloop_b200 import _lib
_lib.SO_PATH = os.path.join(os.path.dirname(_lib.SO_PATH), "libha_b200_probe.so")
from haloop_b200 import ops
dev = torch.device("cuda:0")
def run(B, T, V, U):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, V, generator=g).to(dev)
    tg = torch.randint(1, V, (B, U), generator=g).to(dev)
    il = torch.full((B,), T, dtype=torch.int64, device=dev); tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    for _ in range(2):
        loss, ws = ops.ctc_fwd(x.permute(1, 0, 2), tg, il, tl, True)
    torch.cuda.synchronize()
    pr = ws[:256].view(torch.int64).cpu().tolist()
    for name, o in (("ph1 s101", 0), ("ph1 s102", 8), ("ph2 s101", 16), ("ph2 s102", 24)):
        a = pr[o:o + 5]
        print(f"B={B} T={T} U={U} {name}: fetch {a[1]-a[0]}, advance {a[2]-a[1]}, store/gamma {a[3]-a[2]}, step_end {a[4]-a[3]}, total {a[4]-a[0]}")
    print("  ph1 s101->s102:", pr[8] - pr[0], " ph2 s101->s102:", pr[24] - pr[16])
run(16, 1500, 64, 30)
run(256, 1500, 1024, 300)

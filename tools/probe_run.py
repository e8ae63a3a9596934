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
    for name, o in (("step100", 0), ("step101", 8), ("stepT-50(phase2)", 16)):
        a = pr[o:o + 7]
        print(f"B={B} T={T} U={U} {name}: wait_em {a[1]-a[0]}, decode {a[2]-a[1]}, update {a[3]-a[2]}, store/gamma {a[4]-a[3]}, barrier {a[5]-a[4]}, leader {a[6]-a[5]}, total {a[6]-a[0]}")
    print("  step100 start -> step101 start:", pr[8] - pr[0])
run(16, 1500, 64, 30)
run(256, 1500, 1024, 300)

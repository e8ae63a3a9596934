"""Fused head + CTC against torch float64 on the same GPU (debug / measurement helper, not a test)."""
import sys
import time

import torch
import torch.nn.functional as F

import haloop_b200 as hb


def ref(h, W, b, tg, il, tl, go):
    h64, W64, b64 = (t.double().detach().requires_grad_(True) for t in (h, W, b))
    lp = F.linear(h64, W64, b64).log_softmax(-1).permute(1, 0, 2)
    loss = F.ctc_loss(lp, tg, il, tl, reduction="none")
    (loss * go.double()).sum().backward()
    return loss.detach(), h64.grad, W64.grad, b64.grad


def run(N, T, D, V, S, seed=0, precision="tf32x3", scale=1.0, time_it=False, check=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    h = torch.randn(N, T, D, device="cuda", generator=g)
    W = torch.randn(V, D, device="cuda", generator=g) * (scale / D ** 0.5)
    b = torch.randn(V, device="cuda", generator=g) * 0.1
    tg = torch.randint(1, V, (N, S), device="cuda", generator=g)
    il = torch.randint(max(T // 2, 2 * S + 1), T + 1, (N,), device="cuda", generator=g); il[0] = T
    tl = torch.randint(max(S // 2, 1), S + 1, (N,), device="cuda", generator=g); tl[0] = S
    go = torch.rand(N, device="cuda", generator=g) + 0.5
    hh, WW, bb = h.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    loss = hb.linear_ctc_forward_score(hh, WW, bb, tg, il, tl, precision=precision)
    (loss * go).sum().backward()
    torch.cuda.synchronize()
    if not check:
        print(f'N={N} T={T} D={D} V={V} S={S} {precision}: mean loss {float(loss.mean()):.4f}', flush=True)
    rl, rdh, rdW, rdb = ref(h, W, b, tg, il, tl, go) if check else (None,) * 4
    e = lambda a, r: float((a.double() - r).abs().max())
    if check:
      print(f"N={N} T={T} D={D} V={V} S={S} {precision}: loss rel {float(((loss.double() - rl) / rl).abs().max()):.2e} "
          f"dh {e(hh.grad, rdh):.2e} (max {float(rdh.abs().max()):.2e}) dW {e(WW.grad, rdW):.2e} (max {float(rdW.abs().max()):.2e}) "
          f"db {e(bb.grad, rdb):.2e} (max {float(rdb.abs().max()):.2e})", flush=True)
    if time_it:
        for _ in range(2):
            hh.grad = None
            l2 = hb.linear_ctc_forward_score(hh, WW, bb, tg, il, tl, precision=precision); l2.sum().backward()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        l2 = hb.linear_ctc_forward_score(hh, WW, bb, tg, il, tl, precision=precision)
        ev[1].record()
        l2.sum().backward()
        ev[2].record()
        torch.cuda.synchronize()
        print(f"   fwd {ev[0].elapsed_time(ev[1]):.3f} ms  bwd {ev[1].elapsed_time(ev[2]):.3f} ms", flush=True)
        from haloop_b200 import ops
        loss, saved = ops.head_ctc_fwd(h, W, b, tg, il, tl, ops._PRECISION[precision])
        gout = torch.ones(N, device="cuda")
        for i in range(8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            ops.head_ctc_bwd(h, W, b, saved, gout, S, ops._PRECISION[precision])
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print(f"   bwd op {i}: gpu {e0.elapsed_time(e1):.2f} ms, host call {1e3 * (t1 - t0):.2f} ms", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(3, 70, 64, 40, 7)
        run(3, 70, 64, 40, 7, precision="tf32")
        run(4, 300, 256, 512, 40)
        run(2, 130, 100, 260, 11)
    elif which == "mid":
        run(16, 1000, 1024, 1024, 200, time_it=True)
        run(16, 1000, 1024, 1024, 200, precision="tf32", time_it=True)
        run(16, 1000, 1024, 256, 100, time_it=True)
    elif which == "full":
        run(256, 1500, 1024, 1024, 300, time_it=True)
    elif which == "perf":
        run(256, 1500, 1024, 1024, 300, time_it=True, check=False)
    elif which == "prof":          # one forward + backward of a quarter batch (1 + 3 x 21 GEMM launches; ncu takes the first six)
        run(64, 1500, 1024, 1024, 300, check=False)
    elif which == "perf1":
        run(256, 1500, 1024, 1024, 300, time_it=True, check=False, precision="tf32")

"""Time the GPU prefix beam search at BASELINE-like sizes against the reference's Python (reduced T)."""
import sys, time, torch
sys.path.insert(0, ".")
import haloop_b200 as hb
g = torch.Generator(device="cuda").manual_seed(0)
for (N, T, V, beam) in [(32, 1500, 1024, 3), (32, 1500, 1024, 8), (256, 1500, 1024, 8), (32, 1500, 1024, 16), (32, 500, 256, 8)]:
    x = torch.randn(N, T, V, device="cuda", generator=g)
    x[:, :, 0] += 4.0 * (torch.rand(N, T, device="cuda", generator=g) < 0.7)
    lp = x.log_softmax(-1)
    il = torch.full((N,), T, device="cuda")
    hb.ctc_beam_search_decode_logits(lp, beam, il, graves=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hyp, hl, sc = hb.ctc_beam_search_decode_logits(lp, beam, il, graves=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"N={N} T={T} V={V} beam={beam}: {dt * 1e3:.1f} ms  ({N * T / dt:.3g} frames/s), best hyp len {int(hl[0, 0])}", flush=True)

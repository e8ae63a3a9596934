#!/usr/bin/env python
"""bench.py — loss + logit-gradient throughput of the alignment-loss hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ctc|star|rnnt] [--impl b200|reference]

A "step" is one pass of the hot path (forward loss + backward logit gradient, grad_output = 1/N)
over one synthetic batch.  Workloads are BASELINE.json's configs: ctc = configs[1]
(B=256 T=1500 V=1024 U=300, the headline), star = configs[2], rnnt = configs[3].  With N > 1 every
rank owns a full batch of its own (utterances are independent: weak scaling, no data-path
collective) and the only collective is the all-reduce of [sum loss, count] (configs[4]).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, B, T, V, U)
    "ctc": ("ctc", 256, 1500, 1024, 300),
    "star": ("star", 128, 1000, 512, 200),
    "rnnt": ("rnnt", 32, 500, 1024, 100),
    "ctc_c1": ("ctc", 8, 200, 256, 50),
    # config 2 with a batch that is a multiple of the 148 SMs (the trellis kernel runs one CTA per utterance, two per
    # SM): shows how much of config 2's time is the 256-on-148 imbalance
    "ctc_b296": ("ctc", 296, 1500, 1024, 300),
    # the same RNN-T batch given as its two factors f (N,T,V), g (N,U+1,V) instead of the
    # 6.6 GB joint f[:, :, None, :] + g[:, None, :, :] (ha/recognizer.py:114)
    "rnnt_fg": ("rnnt_fg", 32, 500, 1024, 100),
}


def alg_bytes(kind, B, T, V, U):
    """read the logits once + write the logit gradient once, fp32."""
    if kind == "rnnt_fg":
        return 8 * V * B * (T + U + 1)          # read f and g once, write their gradients once
    return 8 * V * B * T * ((U + 1) if kind == "rnnt" else 1)


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the latest committed ncu --set full summary
    (profiles/traffic_rNN.json, BASELINE-size workloads only); None if there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))
    if not files:
        return None
    try:
        return json.load(open(files[-1])).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled while the timed region runs: NVML in a thread every 10 ms
    (nvidia-smi -lms as the fallback: its first sample takes ~0.3 s, too slow for a 0.2 s region)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x4, "sw_power_cap"), (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.uuid = uuid
        self.rows = []          # nvidia-smi fallback
        self.sm, self.mask = [], 0
        self.proc = None
        self.nvml = None
        self.run = False
        self.prepared = False

    def prepare(self):
        """NVML initialisation takes ~10 ms of host time: done before the pre-timing barrier, so that rank 0
        does not enter the timed region late and make the other ranks wait in their first all-reduce."""
        if self.prepared:
            return
        self.prepared = True
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nvml, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def start(self):
        self.prepare()
        if self.nvml:
            self.run = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while self.run:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.run = False
            self.th.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def make_inputs(kind, B, T, V, U, seed, device=None, pin=False):
    """Synthetic batch, SURVEY.md §8(d): randn logits, labels in 1..V-1, full lengths."""
    import torch
    g = torch.Generator().manual_seed(seed)
    if kind == "rnnt_fg":
        f = torch.randn(B, T, V, generator=g); gg = torch.randn(B, U + 1, V, generator=g)
        if pin:
            f, gg = f.pin_memory(), gg.pin_memory()
        tg = torch.randint(1, V, (B, U), generator=g)
        il = torch.full((B,), T, dtype=torch.int64); tl = torch.full((B,), U, dtype=torch.int64)
        if device is not None:
            f, gg, tg, il, tl = (t.to(device) for t in (f, gg, tg, il, tl))
        return (f, gg), tg, il, tl
    shape = (B, T, V) if kind != "rnnt" else (B, T, U + 1, V)
    x = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
    # filled in chunks: one 6.6 GB randn call is slow and doubles peak host memory
    flat = x.view(-1)
    step = 1 << 26
    for i in range(0, flat.numel(), step):
        flat[i:i + step].normal_(generator=g)
    tg = torch.randint(1, V, (B, U), generator=g)
    il = torch.full((B,), T, dtype=torch.int64)
    tl = torch.full((B,), U, dtype=torch.int64)
    if device is not None:
        x, tg, il, tl = x.to(device), tg.to(device), il.to(device), tl.to(device)
    return x, tg, il, tl


def cpu_reference_run(kind, B, T, V, U, seconds_target, threads):
    """Time the CPU restatement (oracle, kind "port": the reference is pure Python + PyTorch autograd
    and does not travel to the GPU box) on a bounded sample of the workload: same T, V, U, fewer
    utterances.  Returns (frames_per_s, sample_description, B_cpu, seconds)."""
    import numpy as np
    os.environ["OMP_NUM_THREADS"] = str(threads)      # must be set before libgomp initialises
    from oracle import oracle
    oracle.build()
    try:                                               # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every core
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except Exception:
        pass
    rng = np.random.default_rng(0)

    def run(b):
        tg = rng.integers(1, V, (b, U)); il = np.full(b, T); tl = np.full(b, U)
        if kind == "rnnt_fg":
            f = rng.standard_normal((b, T, V), dtype=np.float32); g = rng.standard_normal((b, U + 1, V), dtype=np.float32)
            t0 = time.perf_counter()
            oracle.rnnt_fg(f, g, tg, il, tl)        # builds the float64 joint, as the reference call site does
            return time.perf_counter() - t0
        shape = (T, b, V) if kind != "rnnt" else (b, T, U + 1, V)
        x = rng.standard_normal(shape, dtype=np.float32).astype(np.float64)
        t0 = time.perf_counter()
        if kind == "ctc":
            oracle.ctc(x, tg, il, tl)
        elif kind == "star":
            oracle.star(x, tg, il, tl, star_penalty=-0.5)
        else:
            oracle.rnnt(x, tg, il, tl)
        return time.perf_counter() - t0

    b0 = max(1, min(B, threads))
    if kind in ("rnnt", "rnnt_fg"):
        b0 = max(1, min(B, threads // 2, 4))
    t_probe = run(b0)
    b = int(max(b0, min(B, b0 * seconds_target / max(t_probe, 1e-3))))
    if kind in ("rnnt", "rnnt_fg"):
        b = min(b, 8)            # 0.8 GB of float64 joint per utterance
    b = max(b0, (b // b0) * b0)
    t = run(b) if b != b0 else t_probe
    return b * T / t, f"B_cpu={b} of B={B}, same T={T} V={V} U={U}, float64, loss+grad", b, t


def python_reference_run(kind, T, V, U, b_cpu):
    """Time the REAL reference (oracle/_ref/ha, copied unmodified from /root/reference by oracle/build_ref.py) on
    the host cores: fp32, log_softmax + ha.ctc.ctc_forward_score3 / ha.star.star_ctc_forward_score /
    ha.transducer.transducer_forward_score + .sum().backward(), on `b_cpu` utterances of the same T, V, U
    (the full batches need ~42 / 8 / 7 minutes).  Returns None when oracle/_ref is absent."""
    from oracle import build_ref
    if not build_ref.available() or kind == "rnnt_fg":
        return None
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    try:
        from ha.ctc import ctc_forward_score3
        from ha.star import star_ctc_forward_score
        from ha.transducer import transducer_forward_score
    finally:
        sys.path.pop(0)
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    shape = (T, b_cpu, V) if kind != "rnnt" else (b_cpu, T, U + 1, V)
    x = torch.randn(shape, generator=g, requires_grad=True)
    tg = torch.randint(1, V, (b_cpu, U), generator=g)
    il = torch.full((b_cpu,), T); tl = torch.full((b_cpu,), U)
    t0 = time.perf_counter()
    lp = x.log_softmax(-1)
    if kind == "ctc":
        loss = ctc_forward_score3(lp, tg, il, tl)
    elif kind == "star":
        loss = star_ctc_forward_score(lp, tg, il, tl, star_penalty=-0.5)
    else:
        loss = transducer_forward_score(lp, tg, il, tl)
    loss.sum().backward()
    dt = time.perf_counter() - t0
    return {"value": b_cpu * T / dt, "unit": "frames/s", "cores": os.cpu_count() or 1, "torch_threads": torch.get_num_threads(),
            "kind": "reference", "seconds": dt,
            "sample": f"the reference's own Python (ha/{'transducer' if kind == 'rnnt' else kind}.py, fp32, autograd backward) "
                      f"on B_cpu={b_cpu} utterances of the same T={T} V={V} U={U}"}


def measure_sweep(kind, rank, world, dev, steps, warm, n_streams=None, budget_mb=None):
    """BASELINE.json configs[4] / SURVEY 8(d) C5: a pool of variable-length utterances, sorted by length, cut
    into buckets by padded byte cost (the DurationBatchSampler rule, ha/sampler.py:13-29; RNN-T: in strips of
    similar T, then by U, so that both paddings stay small), dealt to the ranks greedily by cost; a step is one
    pass over the pool (loss + gradient of every bucket) and one all-reduce of [sum loss, count].  Strong
    scaling: the pool is fixed as the number of GPUs grows.  Needs an initialised process group when world > 1.
    Returns the result dict on rank 0 (None elsewhere)."""
    import random
    import torch
    import torch.distributed as dist
    from haloop_b200 import ops, sharding
    V = 1024
    rnd = random.Random(0)
    if kind == "ctc":
        n_pool = 4096
        tl_ = [10 * rnd.randint(20, 150) for _ in range(n_pool)]
        ul_ = [max(1, min((t - 1) // 2, round(t / 5 * rnd.uniform(0.6, 1.0)))) for t in tl_]
        budget = 6_291_456_000                       # four times BASELINE config 2's padded logits: the fused CTC kernels skip
                                                     # padded frames, so large ragged buckets (fewer, fuller launches) beat tight
                                                     # ones: 9.8 ms per pass with 3 buckets against 12.0 ms with 19 (0.79 GB)
    else:
        n_pool = 256                                 # 4096 RNN-T joints (~300 GB) do not fit one GPU: 256 do
        tl_ = [rnd.randint(100, 500) for _ in range(n_pool)]
        ul_ = [rnd.randint(20, 100) for _ in range(n_pool)]
        budget = 827_392_000                         # an eighth of BASELINE config 4's padded joint
    if budget_mb:
        budget = int(budget_mb * 1e6)
    n_streams = n_streams or (2 if kind == "ctc" else 8)
    # utterances are dealt to the ranks by their own cost, then every rank buckets its share by length: the ranks'
    # loads agree to a fraction of a percent and the buckets stay large at any number of ranks
    shares = sharding.deal_utterances(tl_, ul_, V, world, kind)
    all_buckets = []
    for r in range(world):
        bs = sharding.bucket_by_length([tl_[i] for i in shares[r]], [ul_[i] for i in shares[r]], V, budget, kind)
        for b in bs:
            b.indices = [shares[r][i] for i in b.indices]
        all_buckets.append(bs)
    buckets = [b for bs in all_buckets for b in bs]
    mine_b = all_buckets[rank]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    data = []
    for b in mine_b:
        Bk = len(b.indices)
        shape = (Bk, b.t_max, V) if kind == "ctc" else (Bk, b.t_max, b.u_max + 1, V)
        x = torch.randn(shape, device=dev, generator=g)
        il = torch.tensor([tl_[i] for i in b.indices], device=dev)
        tl = torch.tensor([ul_[i] for i in b.indices], device=dev)
        tg = torch.randint(1, V, (Bk, b.u_max), device=dev, generator=g)
        tg = tg * (torch.arange(b.u_max, device=dev)[None, :] < tl[:, None])       # 0-padded (ha/loop.py:40)
        data.append((x, tg, il, tl, torch.ones(Bk, device=dev)))
    red = torch.zeros(2, device=dev, dtype=torch.float64)

    # buckets rotate over the side streams: a short bucket's kernels (one CTA per utterance and sweep direction) do
    # not fill the GPU on their own, the neighbouring buckets' kernels run beside them
    main_stream = torch.cuda.current_stream()
    side = [torch.cuda.Stream() for _ in range(n_streams)]

    def step():
        res = sharding.bucketed_pass(kind, data, side)
        red[0] = sum(l.sum(dtype=torch.float64) for l, _ in res)
        red[1] = float(sum(l.numel() for l, _ in res))
        if world > 1:
            dist.all_reduce(red)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(dev.index or 0, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / steps
    del data
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    frames = sum(tl_)
    if kind == "ctc":
        ab = 8 * V * frames
        padded = sum(len(b.indices) * b.t_max for b in buckets)
        true_units = frames
    else:
        true_units = sum(t * (u + 1) for t, u in zip(tl_, ul_))
        ab = 8 * V * true_units
        padded = sum(len(b.indices) * b.t_max * (b.u_max + 1) for b in buckets)
    peak, peak_src = measured_peak()
    gbs = ab / (ms_step * 1e-3) / 1e9
    loads = [sum(b.cost for b in bs) for bs in all_buckets]
    return {
        "metric": "utterance-frames/sec (loss + logit gradient)", "value": frames / (ms_step * 1e-3),
        "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"sweep_{kind}: {kind} pool of {n_pool} variable-length utterances, V={V}, "
                               f"dealt to {world} rank(s) by cost, each share cut into length buckets of <= "
                               f"{budget / 1e9:.2f} GB padded logits ({len(buckets)} in all); one pass over the pool per step, "
                               f"{n_streams} buckets in flight",
                   "frames": frames, "buckets": len(buckets),
                   "padding_overhead": padded / true_units,
                   "rank_load_imbalance": max(loads) / (sum(loads) / len(loads)),
                   "l2": "every bucket's logits are larger than L2"},
        "mean_loss": float(red[0] / red[1]),
        "gpu_launches": (3 if kind == "ctc" else 5) * len(mine_b) * steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "scope": "whole step (all buckets, all ranks), algorithmic bytes of the TRUE "
                                              "lengths, against world x the single-GPU peak", "achieved": gbs,
                     "peak": peak * world, "unit": "GB/s", "frac": gbs / peak / world, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": ab},
    }


def run_sweep(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kind = "rnnt" if args.workload == "sweep_rnnt" else "ctc"
    out = measure_sweep(kind, rank, world, dev, max(1, min(args.steps, 10)), max(1, min(args.warmup, 3)),
                        args.sweep_streams, args.sweep_budget_mb)
    if rank == 0:
        out["e2e"] = None; out["cpu_baseline"] = None
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def measure_quick(kind, B, T, V, U, dev, scale=1.0, steps=20, warm=3):
    """A short single-GPU measurement of one more workload (device-generated inputs, CUDA events around `steps`
    steps after `warm`): the sub-records of the default run's "workloads"."""
    import torch
    from haloop_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5)
    nbytes_x = alg_bytes(kind, B, T, V, U) // 2
    nsets = 1 if nbytes_x > 2 * 126e6 else max(2, int(4 * 126e6 // nbytes_x) + 1)
    tg = torch.randint(1, V, (B, U), device=dev, generator=g)
    il = torch.full((B,), T, dtype=torch.int64, device=dev); tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    if kind == "rnnt_fg":
        xs = [(torch.randn(B, T, V, device=dev, generator=g) * scale, torch.randn(B, U + 1, V, device=dev, generator=g) * scale)
              for _ in range(nsets)]
    else:
        shape = (B, T, V) if kind != "rnnt" else (B, T, U + 1, V)
        xs = [torch.randn(shape, device=dev, generator=g) * scale for _ in range(nsets)]
    gout = torch.ones(B, device=dev)

    def step(i):
        x = xs[i % nsets]
        if kind == "ctc":
            xv = x.permute(1, 0, 2); loss, ws = ops.ctc_fwd(xv, tg, il, tl, True); ops.ctc_bwd(xv, ws, gout, U, True)
        elif kind == "star":
            xv = x.permute(1, 0, 2); loss, ws = ops.star_fwd(xv, tg, il, tl, -0.5, True); ops.star_bwd(xv, ws, gout, U, True)
        elif kind == "rnnt_fg":
            loss, ws = ops.rnnt_fg_fwd(x[0], x[1], tg, il, tl); ops.rnnt_fg_bwd(x[0], x[1], ws, gout)
        else:
            loss, ws = ops.rnnt_fwd(x, tg, il, tl, True); ops.rnnt_bwd(x, ws, gout, True)
        return loss

    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    peak, _ = measured_peak()
    ab = alg_bytes(kind, B, T, V, U)
    out = {"config": f"{kind} B={B} T={T} V={V} U={U} fp32, logits x{scale:g}, full lengths, from logits",
           "ms_per_step": ms, "value": B * T / (ms * 1e-3), "unit": "frames/s", "steps": steps, "warmup": warm,
           "mean_loss": float(loss.double().mean()),
           "roofline": {"bound": "hbm", "achieved": ab / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": ab / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": ab},
           "clocks": clocks}
    if kind == "rnnt_fg":
        flops = 6.0 * B * T * (U + 1) * V
        out["roofline"]["note"] = "bound by its GEMMs and the lattice sweep, not by HBM"
        out["gemm_tflops_over_whole_step"] = flops / (ms * 1e-3) / 1e12
    del xs
    torch.cuda.empty_cache()
    return out


def measure_head(dev, B=256, T=1500, D=1024, V=1024, U=300, steps=5, warm=2):
    """Linear -> log_softmax -> CTC as one op (haloop_b200.linear_ctc_forward_score, tcgen05 tf32 x 3)
    on BASELINE config 2 with the reference's feat_dim (ha/recognizer.py:38), loss + gradients w.r.t. features, weight
    and bias; beside it the unfused path on the same GPU (cuBLAS fp32 logits, allow_tf32 off as under the reference's
    autocast(float32) -> the fused CTC kernels -> cuBLAS dW, dh).  Tensor-core bound: the roofline is in TFLOP/s."""
    import torch
    import torch.nn.functional as F
    from haloop_b200 import ops
    g = torch.Generator(device=dev).manual_seed(7)
    h = torch.randn(B, T, D, device=dev, generator=g)
    W = torch.randn(V, D, device=dev, generator=g) / D ** 0.5
    b = torch.randn(V, device=dev, generator=g) * 0.1
    tg = torch.randint(1, V, (B, U), device=dev, generator=g)
    il = torch.full((B,), T, dtype=torch.int64, device=dev); tl = torch.full((B,), U, dtype=torch.int64, device=dev)
    gout = torch.ones(B, device=dev)

    def fused(prec):
        loss, saved = ops.head_ctc_fwd(h, W, b, tg, il, tl, prec)
        ops.head_ctc_bwd(h, W, b, saved, gout, U, prec)
        return loss

    def unfused():
        logits = F.linear(h, W, b)
        xv = logits.permute(1, 0, 2)
        loss, ws = ops.ctc_fwd(xv, tg, il, tl, True)
        gl = ops.ctc_bwd(xv, ws, gout, U, True).permute(1, 0, 2).reshape(B * T, V)
        dh = gl @ W; dW = gl.t() @ h.reshape(B * T, D); db = gl.sum(0)       # noqa: F841
        return loss

    def timed(fn):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(loss.double().mean()), (torch.cuda.max_memory_allocated(dev) - base) / 1e9

    sampler = ClockSampler(dev.index or 0, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    sampler.start()
    ms3, loss3, mem3 = timed(lambda: fused(3))
    clocks = sampler.stop()
    ms1, loss1, mem1 = timed(lambda: fused(1))
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        msu, lossu, memu = timed(unfused)
        torch.backends.cuda.matmul.allow_tf32 = True
        msu_tf32, _, _ = timed(unfused)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_tf32, src = float(peaks["bf16_tflops_sustained"]) / 2, "measured bf16_tflops_sustained / 2 (kind::tf32 runs at half the bf16 rate)"
    except Exception:
        peak_tf32, src = 2250.0 / 2, "fallback: nominal dense bf16 2250 TFLOP/s / 2"
    gemm = 2.0 * B * T * D * V                       # one h W^T-sized contraction
    mma3 = 4 * 3 * gemm                              # forward, recomputation, dh, dW; three tf32 products each
    return {"config": f"head_ctc B={B} T={T} D={D} V={V} U={U} fp32 features/weights, full lengths; loss + d/dh, d/dW, d/db",
            "ms_per_step": ms3, "value": B * T / (ms3 * 1e-3), "unit": "frames/s", "steps": steps, "warmup": warm,
            "mean_loss": loss3, "dtype": "tf32 x 3 (fp32-grade), fp32 accumulate", "peak_extra_memory_gb": mem3,
            "roofline": {"bound": "tensor", "achieved": mma3 / (ms3 * 1e-3) / 1e12, "peak": peak_tf32, "unit": "TFLOP/s",
                         "frac": mma3 / (ms3 * 1e-3) / 1e12 / peak_tf32, "peak_source": src,
                         "flops_per_step": mma3, "note": "tf32 tensor-core flops actually issued (4 contractions x 3 split "
                         "products); the useful fp32 flops are a third of it"},
            "single_product_tf32": {"ms_per_step": ms1, "mean_loss": loss1},
            "unfused_same_gpu": {"what": "F.linear (cuBLAS fp32) -> ha_ctc_fwd/bwd -> cuBLAS dh, dW, db", "ms_per_step": msu,
                                 "ms_per_step_allow_tf32": msu_tf32, "mean_loss": lossu, "peak_extra_memory_gb": memu},
            "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="ctc", choices=sorted(WORKLOADS) + ["sweep_ctc", "sweep_rnnt"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--graph", action="store_true",
                    help="replay the step from a CUDA graph (launch-bound small workloads such as ctc_c1); the "
                         "rotating inputs are copied into the captured buffer every step")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="skip timing F.ctc_loss / torchaudio.rnnt_loss on the same GPU (SURVEY 8d)")
    ap.add_argument("--sweep-streams", type=int, default=None, help="sweep workloads: buckets in flight")
    ap.add_argument("--sweep-budget-mb", type=float, default=None, help="sweep workloads: padded logit bytes per bucket")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline workload only: skip the \"workloads\" sub-records (star, rnnt, rnnt_fg, ctc with x3 "
                         "logits and the length-bucketed sweeps; with N > 1 the sweeps are the strong-scaling curve)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload.startswith("sweep"):
        return run_sweep(args, rank, world, local_rank)
    kind, B, T, V, U = WORKLOADS[args.workload]
    metric = "utterance-frames/sec (loss + logit gradient)"
    config = {"workload": f"{args.workload}: {kind} B={B}/GPU T={T} V={V} U={U} fp32, full lengths, from logits",
              "batch_per_gpu": B, "T": T, "V": V, "U": U, "sharding": f"batch x{max(world, args.gpus)} (utterances independent)",
              "l2": "inputs larger than L2 (no flush needed)" if alg_bytes(kind, B, T, V, U) // 2 > 2 * 126e6
              else "inputs rotate over 4 buffer sets larger than L2 in total"}
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU implementation of the path, all host threads, bounded sample per step
        vals = []
        per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
        for i in range(args.warmup + args.steps):
            fps, sample, b, t = cpu_reference_run(kind, B, T, V, U, per_step, threads)
            if i >= args.warmup:
                vals.append((fps, t))
        v = sum(f for f, _ in vals) / len(vals)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t for _, t in vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    import haloop_b200 as hb
    from haloop_b200 import ops

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # small workloads rotate over several input sets so that no step finds its logits in L2
    nbytes_x = alg_bytes(kind, B, T, V, U) // 2
    nsets = 1 if nbytes_x > 2 * 126e6 else max(2, int(4 * 126e6 // nbytes_x) + 1)
    sets = [make_inputs(kind, B, T, V, U, seed=1000 * rank + s, device=dev) for s in range(nsets)]
    gout = torch.full((B,), 1.0, device=dev)
    # [sum loss, count] of the last two steps: the all-reduce of step k is only waited for when step k+2 reuses
    # its buffer, so the collective (the loss is needed for logging, not by the data path) overlaps the kernels
    reds = [torch.zeros(2, device=dev, dtype=torch.float64) for _ in range(2)]
    pending = [None, None]

    def view(x):
        # the reference's call site hands the loss logits.permute(1,0,2) of an (N,T,C) buffer
        return x.permute(1, 0, 2) if kind in ("ctc", "star") else x

    fwd_ev, bwd_ev = [], []

    def step(i, timed=False):
        x, tg, il, tl = sets[i % nsets]
        xv = view(x)
        if timed:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        if kind == "ctc":
            loss, ws = ops.ctc_fwd(xv, tg, il, tl, True)
        elif kind == "star":
            loss, ws = ops.star_fwd(xv, tg, il, tl, -0.5, True)
        elif kind == "rnnt_fg":
            loss, ws = ops.rnnt_fg_fwd(xv[0], xv[1], tg, il, tl)
        else:
            loss, ws = ops.rnnt_fwd(xv, tg, il, tl, True)
        if timed:
            e1.record()
        if kind == "ctc":
            gx = ops.ctc_bwd(xv, ws, gout, U, True)
        elif kind == "star":
            gx = ops.star_bwd(xv, ws, gout, U, True)
        elif kind == "rnnt_fg":
            gx = ops.rnnt_fg_bwd(xv[0], xv[1], ws, gout)
        else:
            gx = ops.rnnt_bwd(xv, ws, gout, True)
        if timed:
            e2.record()
            fwd_ev.append((e0, e1)); bwd_ev.append((e1, e2))
        red = reds[i & 1]
        if pending[i & 1] is not None:
            pending[i & 1].wait()
        red[0] = loss.sum(); red[1].fill_(B)
        if world > 1:
            pending[i & 1] = dist.all_reduce(red, async_op=True)   # the path's only collective: [sum loss, count]
        return gx

    if args.graph and kind in ("ctc", "star", "rnnt") and world == 1:
        eager_step = step
        xbuf = sets[0][0].clone()
        _, tg0, il0, tl0 = sets[0]
        gstream = torch.cuda.Stream()
        gstream.wait_stream(torch.cuda.current_stream())

        def body():
            xv = view(xbuf)
            if kind == "ctc":
                loss, ws = ops.ctc_fwd(xv, tg0, il0, tl0, True); gx = ops.ctc_bwd(xv, ws, gout, U, True)
            elif kind == "star":
                loss, ws = ops.star_fwd(xv, tg0, il0, tl0, -0.5, True); gx = ops.star_bwd(xv, ws, gout, U, True)
            else:
                loss, ws = ops.rnnt_fwd(xv, tg0, il0, tl0, True); gx = ops.rnnt_bwd(xv, ws, gout, True)
            reds[0][0] = loss.sum(); reds[0][1].fill_(B)
            return gx
        with torch.cuda.stream(gstream):
            body()
        torch.cuda.current_stream().wait_stream(gstream)
        cgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cgraph):
            body()

        def step(i, timed=False):           # noqa: F811  (same targets/lengths in every set: only the logits rotate)
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            xbuf.copy_(sets[i % nsets][0])
            cgraph.replay()
            if timed:
                e1.record()
                fwd_ev.append((e0, e1)); bwd_ev.append((e1, e1))
        reds[1] = reds[0]
        config["cuda_graph"] = "the step (4-5 kernels + loss sum) is captured once and replayed"
    for i in range(args.warmup):
        step(i)
    try:
        dev_uuid = torch.cuda.get_device_properties(local_rank).uuid
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    if rank == 0:
        sampler.prepare()
    barrier()
    if rank == 0:
        sampler.start()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step(i, timed=True)
    for h in pending:                       # the last two all-reduces complete inside the timed region
        if h is not None:
            h.wait()
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    red = reds[(args.steps - 1) & 1]
    mean_loss = float(red[0] / red[1])
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    frames = B * T * world
    value = frames / (ms_step * 1e-3)
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_ev) / len(fwd_ev)
    bwd_ms = sum(a.elapsed_time(b) for a, b in bwd_ev) / len(bwd_ev)

    # ---- end to end: pinned host logits -> device, loss + gradient, loss back to the host, every step;
    #      the copy of step k+1 overlaps the kernels of step k on a second stream
    e2e = None
    if not args.no_e2e:
        hx, htg, hil, htl = make_inputs(kind, B, T, V, U, seed=77 + rank, pin=True)
        htg, hil, htl = htg.pin_memory(), hil.pin_memory(), htl.pin_memory()
        hloss = torch.empty(B, dtype=torch.float32).pin_memory()
        fg = kind == "rnnt_fg"
        dbuf = [tuple(torch.empty_like(t) for t in sets[0][0]) if fg else torch.empty_like(sets[0][0]) for _ in range(2)]
        dtg = [torch.empty_like(sets[0][1]) for _ in range(2)]
        dil = [torch.empty_like(sets[0][2]) for _ in range(2)]
        dtl = [torch.empty_like(sets[0][3]) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream()

        def upload(k):
            s = k & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                if fg:
                    dbuf[s][0].copy_(hx[0], non_blocking=True); dbuf[s][1].copy_(hx[1], non_blocking=True)
                else:
                    dbuf[s].copy_(hx, non_blocking=True)
                dtg[s].copy_(htg, non_blocking=True)
                dil[s].copy_(hil, non_blocking=True); dtl[s].copy_(htl, non_blocking=True)
                ready[s].record(copy_stream)

        def compute(s):
            main_stream.wait_event(ready[s])
            xd = tuple(t.requires_grad_(True) for t in dbuf[s]) if fg else view(dbuf[s]).requires_grad_(True)
            if fg:
                loss = hb.transducer_forward_score_fg(xd[0], xd[1], dtg[s], dil[s], dtl[s])
            elif kind == "ctc":
                loss = hb.ctc_forward_score3(xd, dtg[s], dil[s], dtl[s], from_logits=True)
            elif kind == "star":
                loss = hb.star_ctc_forward_score(xd, dtg[s], dil[s], dtl[s], star_penalty=-0.5, from_logits=True)
            else:
                loss = hb.transducer_forward_score(xd, dtg[s], dil[s], dtl[s], from_logits=True)
            loss.sum().backward()
            hloss.copy_(loss.detach(), non_blocking=True)
            freed[s].record(main_stream)
            for t in (dbuf[s] if fg else (dbuf[s],)):
                t.requires_grad_(False); t.grad = None

        for s in range(2):
            freed[s].record(main_stream)
        n_e2e = max(3, min(args.steps, 8))
        upload(0)
        compute(0)                              # warm-up step (untimed)
        barrier()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record()
        upload(1)
        for k in range(1, 1 + n_e2e):
            if k < n_e2e:
                upload(k + 1)                   # next step's copy overlaps this step's kernels
            compute(k & 1)
        ee.record()
        barrier()
        t2 = torch.tensor([es.elapsed_time(ee)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2) / n_e2e
        h2d = sum(t.numel() for t in (hx if fg else (hx,))) * 4 + htg.numel() * 8 + 16 * B
        e2e = {"value": frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4 * B, "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_gbs": h2d / (e2e_ms * 1e-3) / 1e9,
               "note": "pinned host logits -> device every step (double-buffered on a copy stream), per-utterance loss "
                       "read back; the logit gradient stays on the device, where the model's backward consumes it. "
                       "This number is bound by the host-to-device copy (h2d_gbs, PCIe), not by the kernels"}

    # ---- the other workloads of BASELINE.json, as sub-records of the same line.  One GPU: configs 3, 4, the
    #      joint-free RNN-T, config 2 with x3 logits (peaky posteriors) and both length-bucketed sweeps (config 5).
    #      N GPUs: the sweeps, which are the STRONG-scaling curve of config 5 (the headline above is weak scaling).
    extras = None
    if not args.no_extras and args.workload == "ctc" and not args.graph:
        sets = sets[:1]                 # the library baseline below still needs one input set
        torch.cuda.empty_cache()
        extras = {}
        try:
            if world == 1:
                for name, (k2, B2, T2, V2, U2) in (("star", WORKLOADS["star"]), ("rnnt", WORKLOADS["rnnt"]),
                                                    ("rnnt_fg", WORKLOADS["rnnt_fg"])):
                    extras[name] = measure_quick(k2, B2, T2, V2, U2, dev)
                extras["ctc_x3"] = measure_quick("ctc", B, T, V, U, dev, scale=3.0)
                try:
                    extras["head_ctc"] = measure_head(dev)
                except Exception as e:
                    extras["head_ctc"] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            for k2 in ("ctc", "rnnt"):
                r = measure_sweep(k2, rank, world, dev, 5, 2)
                if rank == 0:
                    extras["sweep_" + k2] = r
        except Exception as e:                      # a sub-record must never take the headline down with it
            extras["error"] = repr(e)[:300]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    ab = alg_bytes(kind, B, T, V, U)
    mid = "lattice" if kind.startswith("rnnt") else "trellis"
    names = [f"{kind}_rows_kernel", f"{kind}_{mid}_kernel", f"{kind}_grad_kernel"]
    if kind == "ctc":
        # fused path (csrc/ctc2.cuh): the forward call is ctc2_prep + ctc2_fwd (row statistics, emission gather and
        # the first half of both sweeps in one kernel), the backward call is ctc2_bwd (second half of the sweeps,
        # occupancies and the gradient rows in one kernel); emissions and occupancies never reach HBM
        names = ["ctc2_prep_kernel", "ctc2_fwd_kernel", "ctc2_bwd_kernel"]
    if kind == "star":
        # the same two-kernel shape for star-CTC (csrc/star2.cuh, round 2): quads (blank, star, blank, label) per position,
        # the dense-over-V gradient formed in the row ring
        names = ["ctc2_prep_kernel", "star2_fwd_kernel", "star2_bwd_kernel"]
    if kind == "rnnt_fg":
        names = ["rnnt_fg_stats_kernel + rnnt_fg_gemm_kernel<E>", "rnnt_lattice_kernel",
                 "rnnt_fg_gemm_kernel<DF> + rnnt_fg_gemm_kernel<DG> + rnnt_fg_fix_kernel"]
    tr = [ncu_traffic(k) if args.workload in ("ctc", "star", "rnnt") else None for k in names]
    launches_per_step = {"ctc": 3, "star": 3, "rnnt": 5, "rnnt_fg": 8}[kind]
    step_gbs = ab / (ms_step * 1e-3) / 1e9
    grad_gbs = ab / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else None      # --graph: forward and backward are one replay
    out = {
        "metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "mean_loss": mean_loss, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        # the whole step against the HBM roofline on ALGORITHMIC bytes (read the logits once + write the gradient
        # once): the conservative figure.  "kernels" splits it: the backward call is one streaming
        # kernel with exactly those algorithmic bytes; the forward call is rows + trellis/lattice.
        "roofline": {
            "bound": "hbm", "scope": "whole step: " + " + ".join(names),
            "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
            "traffic": (sum(t for t in tr if t is not None) if tr[1] is not None and tr[2] is not None else None),
            "peak_source": peak_src, "algorithmic_bytes_per_step": ab,
            "kernels": {
                names[2]: {"ms": bwd_ms, "achieved": grad_gbs, "frac": (grad_gbs / peak) if grad_gbs else None, "traffic": tr[2],
                           "note": "the backward call = this one kernel (timed live by CUDA events): reads the logits, "
                                   "writes the gradient; its algorithmic bytes are the step's"
                                   + ("; it also runs the second half of both sweeps" if kind in ("ctc", "star") else "")},
                "forward (" + names[0] + " + " + names[1] + ")": {
                    "ms": fwd_ms, "traffic": ((tr[0] or 0) + tr[1]) if tr[1] is not None else None,
                    "note": ("one fused kernel: logit rows in through TMA, row statistics + emission gather by the row "
                             "warps, first half of the alpha and beta sweeps by the trellis warps (profiles/, DESIGN.md)"
                             if kind in ("ctc", "star") else
                             "rows kernel is HBM-bound; the " + mid + " kernel is instruction-issue / latency bound "
                             "(profiles/, DESIGN.md section 5)")},
            },
        },
    }
    if kind == "rnnt_fg":
        # not an HBM-bound path any more: three fp32 GEMMs (6 N T (U+1) V flop) around the latency-bound lattice
        flops = 6.0 * B * T * (U + 1) * V
        out["roofline"]["note"] = ("joint-free RNN-T moves 8 V (T+U+1) bytes per utterance instead of 8 V T (U+1): the "
                                   "step is bound by its fp32 SIMT GEMMs and the lattice sweep, not by HBM")
        out["roofline"]["kernels"] = {"fp32 GEMMs (E = F G^T, W G, W^T F)": {
            "flop_per_step": flops, "achieved_tflops_over_whole_step": flops / (ms_step * 1e-3) / 1e12,
            "peak_tflops_fp32_simt_nominal": 148 * 128 * 2 * 1.965e9 / 1e12}}
        out["joint_path_equivalent"] = {"note": "the same batch through the materialised joint is --workload rnnt",
                                        "joint_bytes": 4 * B * T * (U + 1) * V}
    if e2e:
        out["e2e"] = e2e
    if extras is not None:
        out["workloads"] = extras
    if not args.no_library_baseline and world == 1 and kind == "rnnt_fg":
        try:
            from torchaudio.functional import rnnt_loss
            (f0, g0), tg, il, tl = sets[0]
            fl = f0.detach().clone().requires_grad_(True); gl = g0.detach().clone().requires_grad_(True)
            tg32, il32, tl32 = tg.int(), il.int(), tl.int()

            def lib():       # the reference call site: broadcast joint (ha/recognizer.py:114) + rnnt_loss (:121-126)
                return rnnt_loss(fl[:, :, None, :] + gl[:, None, :, :], tg32, il32, tl32, blank=0, reduction="sum",
                                 fused_log_softmax=True)
            for _ in range(2):
                fl.grad = None; gl.grad = None
                lib().backward()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                fl.grad = None; gl.grad = None
                lib().backward()
            b_.record()
            torch.cuda.synchronize()
            lib_ms = a.elapsed_time(b_) / 3
            out["library_baseline"] = {"name": "f[:, :, None, :] + g[:, None, :, :] -> torchaudio.functional.rnnt_loss",
                                       "value": B * T / (lib_ms * 1e-3), "unit": "frames/s", "ms_per_step": lib_ms,
                                       "torch": torch.__version__}
        except Exception as e:
            out["library_baseline"] = {"unavailable": repr(e)[:200]}
    if not args.no_library_baseline and world == 1 and kind in ("ctc", "rnnt"):
        # the strongest same-box library baselines: log_softmax + F.ctc_loss (reduction='sum') and
        # torchaudio's fused rnnt_loss on the same inputs, forward + backward, CUDA events
        try:
            x, tg, il, tl = sets[0]
            xl = x.detach().clone().requires_grad_(True)
            if kind == "ctc":
                def lib():
                    lp = torch.nn.functional.log_softmax(view(xl), dim=-1)
                    return torch.nn.functional.ctc_loss(lp, tg, il, tl, blank=0, reduction="sum", zero_infinity=False)
                name = "torch.nn.functional.log_softmax + ctc_loss"
            else:
                from torchaudio.functional import rnnt_loss
                tg32, il32, tl32 = tg.int(), il.int(), tl.int()
                def lib():
                    return rnnt_loss(xl, tg32, il32, tl32, blank=0, reduction="sum", fused_log_softmax=True)
                name = "torchaudio.functional.rnnt_loss"
            for _ in range(2):
                xl.grad = None
                lib().backward()
            torch.cuda.synchronize()
            n_lib = 5
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_lib):
                xl.grad = None
                lib().backward()
            b_.record()
            torch.cuda.synchronize()
            lib_ms = a.elapsed_time(b_) / n_lib
            out["library_baseline"] = {"name": name, "value": B * T / (lib_ms * 1e-3), "unit": "frames/s",
                                       "ms_per_step": lib_ms, "torch": torch.__version__}
            del xl
        except Exception as e:                      # library not importable / out of memory: report, do not fail
            out["library_baseline"] = {"unavailable": repr(e)[:200]}
    if not args.no_cpu_baseline and world == 1:
        fps, sample, b, tcpu = cpu_reference_run(kind, B, T, V, U, args.cpu_seconds, threads)
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": sample, "seconds": tcpu}
        try:
            ref = python_reference_run(kind, T, V, U, {"ctc": 2, "star": 4, "rnnt": 1}.get(kind, 1))
        except Exception as e:
            ref = {"unavailable": repr(e)[:200]}
        if ref is not None:
            # the C port above is the conservative CPU arm (it is ~1000x faster than the code it restates);
            # this is the reference itself, same box, same run
            out["cpu_baseline_reference"] = ref
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-side logic of the multi-GPU path: bucketing, dealing, and the loss
all-reduce on a world_size-2 gloo group (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from haloop_b200 import sharding


def test_bucket_rule_matches_duration_batch_sampler():
    """Same growth rule as ha/sampler.py:13-29 with bytes for seconds, on length-sorted input."""
    g = torch.Generator().manual_seed(0)
    T = torch.randint(20, 150, (200,), generator=g).mul(10).tolist()
    U = [max(1, min(t // 5, (t - 1) // 2)) for t in T]
    budget = 8 * 1500 * 1024 * 4
    buckets = sharding.bucket_by_length(T, U, 1024, budget, "ctc")
    seen = sorted(i for b in buckets for i in b.indices)
    assert seen == list(range(200)), "every utterance lands in exactly one bucket"
    for b in buckets:
        assert b.t_max == max(T[i] for i in b.indices) and b.u_max == max(U[i] for i in b.indices)
        assert b.cost <= budget or len(b.indices) == 1
        assert b.cost == len(b.indices) * b.t_max * 1024 * 4
    # sorted by length: buckets are contiguous in T, so padding waste is small
    waste = sum(b.cost for b in buckets) / (sum(T) * 1024 * 4)
    assert waste < 1.15
    rn = sharding.bucket_by_length(T, U, 1024, 32 * 500 * 101 * 1024 * 4, "rnnt")
    assert all(b.cost == len(b.indices) * b.t_max * (b.u_max + 1) * 1024 * 4 for b in rn)


def test_deal_is_balanced_and_deterministic():
    g = torch.Generator().manual_seed(1)
    T = torch.randint(200, 1500, (4096,), generator=g).tolist()
    U = [max(1, t // 6) for t in T]
    buckets = sharding.bucket_by_length(T, U, 1024, 256 * 1500 * 1024 * 4 // 8, "ctc")
    for world in (2, 4, 8):
        owned = sharding.deal_buckets(buckets, world)
        assert sorted(b for o in owned for b in o) == list(range(len(buckets)))
        loads = [sum(buckets[b].cost for b in o) for o in owned]
        assert max(loads) / (sum(loads) / world) < 1.10, "greedy deal keeps ranks within 10% (75 buckets)"
        assert owned == sharding.deal_buckets(buckets, world)


def test_shard_batch_partitions():
    for n, world in ((256, 8), (13, 4), (3, 8)):
        parts = [list(sharding.shard_batch(n, r, world)) for r in range(world)]
        assert sum(parts, []) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        losses = torch.rand(10, generator=g, dtype=torch.float64) * 100
        tl = torch.randint(1, 30, (10,), generator=g).double()
        mine = sharding.shard_batch(10, rank, world)
        loc = losses[mine.start:mine.stop].clone().requires_grad_(True)
        tot = sharding.reduce_loss(loc, 1.0 / tl[mine.start:mine.stop])
        tot.backward()
        ref = (losses / tl).mean()                      # ctc_reduce_mean over the global batch, ha/ctc.py:177-178
        plain = sharding.reduce_loss(losses[mine.start:mine.stop])
        ok = (abs(float(tot) - float(ref)) < 1e-12 and
              torch.allclose(loc.grad, (1.0 / tl[mine.start:mine.stop]) / 10) and
              abs(float(plain) - float(losses.mean())) < 1e-12)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_loss_allreduce_world2_gloo():
    """The path's one collective: [sum loss*w, count] all-reduced; every rank gets the global mean and
    the gradient of its own utterances only."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}


def test_reduce_loss_without_process_group():
    l = torch.tensor([2.0, 4.0], requires_grad=True)
    assert float(sharding.reduce_loss(l)) == 3.0
    assert abs(float(sharding.reduce_loss(l, torch.tensor([0.5, 0.25]))) - 1.0) < 1e-7


def test_vmap_fold_helper_merges_the_vmapped_dimension():
    """ops._fold: the batching rule's only tensor manipulation (CPU tensors suffice)"""
    import torch
    from haloop_b200 import ops
    x = torch.arange(2 * 3 * 4 * 5.0).reshape(3, 2, 4, 5)           # vmap dim 0 (B=3) of (T=2, N=4, V=5)
    f = ops._fold(x, 0, 3, 1)                                        # -> (T, B*N, V)
    assert f.shape == (2, 12, 5)
    for b in range(3):
        assert torch.equal(f[:, 4 * b:4 * (b + 1)], x[b])
    y = torch.arange(4.0)                                            # not batched: broadcast
    g = ops._fold(y, None, 3, 0)
    assert g.shape == (12,) and torch.equal(g, y.repeat(3))
    assert ops._fold(None, None, 3, 0) is None


def test_rnnt_buckets_cost_and_deal():
    from haloop_b200 import sharding
    import random
    r = random.Random(1)
    tl = [r.randint(100, 500) for _ in range(64)]; ul = [r.randint(20, 100) for _ in range(64)]
    b = sharding.bucket_by_length(tl, ul, 1024, 1_000_000_000, "rnnt")
    assert sorted(i for x in b for i in x.indices) == list(range(64))
    for x in b:
        assert x.t_max == max(tl[i] for i in x.indices) and x.u_max == max(ul[i] for i in x.indices)
        assert x.cost == len(x.indices) * x.t_max * (x.u_max + 1) * 1024 * 4
        assert x.cost <= 1_000_000_000 or len(x.indices) == 1
    owned = sharding.deal_buckets(b, 4)
    assert sorted(k for o in owned for k in o) == list(range(len(b)))


def test_deal_utterances_then_bucket_per_rank():
    """bench.py's config-5 dealing: utterances go to ranks by their own cost (loads agree to a fraction of a percent),
    every rank buckets its share; the 2-D order keeps RNN-T padding small."""
    import random
    r = random.Random(0)
    tl = [10 * r.randint(20, 150) for _ in range(4096)]
    ul = [max(1, min((t - 1) // 2, round(t / 5 * r.uniform(0.6, 1.0)))) for t in tl]
    for world in (1, 2, 4, 8):
        shares = sharding.deal_utterances(tl, ul, 1024, world, "ctc")
        assert sorted(i for s in shares for i in s) == list(range(4096))
        loads = [sum(tl[i] for i in s) for s in shares]
        assert max(loads) / (sum(loads) / world) < 1.002
        assert shares == sharding.deal_utterances(tl, ul, 1024, world, "ctc")
    tl = [r.randint(100, 500) for _ in range(256)]; ul = [r.randint(20, 100) for _ in range(256)]
    b1 = sharding.bucket_by_length(tl, ul, 1024, 300_000_000, "rnnt")
    true = sum(t * (u + 1) for t, u in zip(tl, ul))
    assert sum(len(b.indices) * b.t_max * (b.u_max + 1) for b in b1) / true < 1.15
    assert sorted(i for b in b1 for i in b.indices) == list(range(256))


class _ToyAudio(torch.utils.data.Dataset):
    def __init__(self, n):
        g = torch.Generator().manual_seed(3)
        self.d = (torch.rand(n, generator=g) * 14 + 1).tolist()

    def __len__(self):
        return len(self.d)

    def duration(self, i):
        return self.d[i]

    def __getitem__(self, i):
        return i, torch.zeros(int(self.d[i] * 10), 4), "x"


def test_length_bucket_batch_sampler_is_a_duration_batch_sampler_dropin():
    """Same interface and growth rule as ha/sampler.py:7-29, usable as DataLoader(batch_sampler=...) in
    ha/loop.py:502-509; length-sorted, sharded over ranks, same number of batches on every rank."""
    ds = _ToyAudio(500)
    seen, counts = [], []
    for rank in range(4):
        s = sharding.LengthBucketBatchSampler(ds, max_duration=120, rank=rank, world_size=4, seed=1)
        batches = list(s)
        counts.append(len(batches))
        assert len(batches) == len(s)
        for b in batches:
            assert len(b) * max(ds.duration(i) for i in b) <= 120 or len(b) == 1
        seen += [i for b in batches for i in b]
        s.set_epoch(1)
        assert sorted(map(tuple, s)) == sorted(map(tuple, batches)) and list(s) != batches, "reshuffled per epoch"
    assert len(set(counts)) == 1, "DDP ranks take the same number of steps"
    full = list(sharding.LengthBucketBatchSampler(ds, max_duration=120, shuffle=False))
    padded = sum(len(b) * max(ds.duration(i) for i in b) for b in full)
    assert padded / sum(ds.duration(i) for b in full for i in b) < 1.15, "length-sorted: little padding inside a batch"
    assert len(seen) == len(set(seen)) and len(seen) > 0.93 * 500, "ranks own disjoint utterances; at most the tail is dropped"
    loader = torch.utils.data.DataLoader(ds, batch_sampler=sharding.LengthBucketBatchSampler(ds, max_duration=120),
                                         collate_fn=lambda b: (torch.tensor([x[0] for x in b]),
                                                               torch.nn.utils.rnn.pad_sequence([x[1] for x in b], batch_first=True)))
    idx, x = next(iter(loader))
    assert x.shape[0] == len(idx) and x.shape[2] == 4


# ------------------------------------------------------------- hac loop wiring (haloop_b200/loop.py) ---
class _ToySet(torch.utils.data.Dataset):
    """Stand-in for the reference's datasets: items + duration(i) (ha/sampler.py:17)."""
    def __init__(self, n):
        g = torch.Generator().manual_seed(3)
        self.len = torch.randint(5, 40, (n,), generator=g).tolist()
        self.x = [torch.randn(l, 6, generator=g) for l in self.len]
        self.y = [torch.randint(1, 5, (max(1, l // 5),), generator=g) for l in self.len]
    def __len__(self): return len(self.len)
    def duration(self, i): return float(self.len[i])
    def __getitem__(self, i): return self.x[i], self.y[i]


def _collate(items):
    xs, ys = zip(*items)
    il = torch.tensor([len(x) for x in xs]); tl = torch.tensor([len(y) for y in ys])
    return (torch.nn.utils.rnn.pad_sequence(xs, batch_first=True), torch.nn.utils.rnn.pad_sequence(ys, batch_first=True), il, tl)


class _ToySystem:
    """encoder + recognizer called through __call__, as ha/loop.py:125-134 does."""
    def __init__(self):
        torch.manual_seed(0)
        self.encoder = torch.nn.Linear(6, 8)
        self.recognizer = torch.nn.Linear(8, 5)
    def forward(self, x, y, il, tl):
        lp = self.recognizer(torch.tanh(self.encoder(x))).log_softmax(-1).permute(1, 0, 2)
        return torch.nn.functional.ctc_loss(lp, y, il, tl)


def _loop_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from haloop_b200 import loop
        ds = _ToySet(64)
        system = loop.distribute(_ToySystem())
        loader = loop.sharded_loader(ds, _collate, max_duration=200, seed=1)
        loader.batch_sampler.set_epoch(0)
        seen, nb = [], 0
        for x, y, il, tl in loader:
            system.encoder.zero_grad(); system.recognizer.zero_grad()
            system.forward(x, y, il, tl).backward()
            nb += 1
            seen.append(int(il.sum()))
        enc, rec = loop.state_modules(system)
        grads = torch.cat([p.grad.flatten() for p in list(enc.parameters()) + list(rec.parameters())])
        gathered = [torch.zeros_like(grads) for _ in range(world)]
        dist.all_gather(gathered, grads)
        idx = sorted(i for b in loader.batch_sampler for i in b)
        all_idx = [None] * world
        dist.all_gather_object(all_idx, idx)
        nbs = [None] * world
        dist.all_gather_object(nbs, nb)
        out[rank] = dict(same_grads=bool(torch.allclose(gathered[0], gathered[1])), nb=nbs,
                         disjoint=len(set(all_idx[0]) & set(all_idx[1])) == 0, covered=len(set(all_idx[0]) | set(all_idx[1])),
                         wrapped=isinstance(system.encoder, torch.nn.parallel.DistributedDataParallel))
    finally:
        dist.destroy_process_group()


def test_loop_distribute_and_sharded_loader_world2_gloo():
    """DDP around encoder + recognizer and the rank-sharded length-bucketed loader: both ranks step the same number of
    batches over disjoint utterances and end every step with identical (all-reduced) gradients."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_loop_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    for r in range(world):
        o = out[r]
        assert o["wrapped"] and o["same_grads"] and o["disjoint"]
        assert o["nb"][0] == o["nb"][1] and o["nb"][0] >= 2
        assert o["covered"] >= 56            # at most the leftover batches of the longer share are dropped

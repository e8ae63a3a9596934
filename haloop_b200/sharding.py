"""Length bucketing and batch sharding for multi-GPU runs (BASELINE.json configs[4]).

Utterances are independent in all three losses, so the path shards by whole
utterances with no data-path collective: logits and gradients never leave the rank that owns them.
The only exchange is one all-reduce of [sum_b loss_b * w_b, sum_b w_b] per step.

bucket_by_length mirrors the reference's DurationBatchSampler (ha/sampler.py:7-29: grow a batch until
(len(batch)+1) * max_duration would exceed the budget), with padded bytes instead of seconds, after
sorting by length so that padding inside a bucket stays small.
"""
from dataclasses import dataclass
from typing import List, Sequence

import torch
import torch.utils.data


@dataclass
class Bucket:
    indices: List[int]      # utterance ids
    t_max: int
    u_max: int
    cost: int               # padded logit bytes


def padded_cost(n, t_max, u_max, vocab, kind):
    """Bytes of the padded fp32 logits of a bucket: B*T*V (ctc/star) or B*T*(U+1)*V (rnnt)."""
    per = t_max * vocab * 4 * ((u_max + 1) if kind == "rnnt" else 1)
    return n * per


def _length_order(in_lens, tgt_lens, vocab, budget_bytes, kind):
    """CTC / star: by input length.  RNN-T pads BOTH T and U+1, so sorting by T alone leaves U random inside a
    bucket (1.6x padded lattice nodes on a uniform pool): cut the T-sorted list into strips of about
    sqrt(utterances per bucket) buckets each and sort every strip by U, so a bucket is compact in both."""
    order = sorted(range(len(in_lens)), key=lambda i: (in_lens[i], tgt_lens[i]))
    if kind != "rnnt" or len(order) < 4:
        return order
    mean_cost = sum(padded_cost(1, in_lens[i], tgt_lens[i], vocab, kind) for i in order) / len(order)
    per_bucket = max(1.0, budget_bytes / mean_cost)
    n_buckets = max(1.0, len(order) / per_bucket)
    strips = max(1, round(n_buckets ** 0.5))
    size = -(-len(order) // strips)
    out = []
    for k in range(0, len(order), size):
        out += sorted(order[k:k + size], key=lambda i: (tgt_lens[i], in_lens[i]))
    return out


def bucket_by_length(in_lens: Sequence[int], tgt_lens: Sequence[int], vocab: int, budget_bytes: int,
                     kind: str = "ctc") -> List[Bucket]:
    """Sort by length (see _length_order), then cut greedily: a bucket closes when adding the next utterance
    would push its padded cost past `budget_bytes` (the DurationBatchSampler rule, ha/sampler.py:13-29)."""
    order = _length_order(in_lens, tgt_lens, vocab, budget_bytes, kind)
    buckets, cur, t_max, u_max = [], [], 0, 0
    for i in order:
        nt, nu = max(t_max, in_lens[i]), max(u_max, tgt_lens[i])
        if cur and padded_cost(len(cur) + 1, nt, nu, vocab, kind) > budget_bytes:
            buckets.append(Bucket(cur, t_max, u_max, padded_cost(len(cur), t_max, u_max, vocab, kind)))
            cur, nt, nu = [], in_lens[i], tgt_lens[i]
        cur.append(i)
        t_max, u_max = nt, nu
    if cur:
        buckets.append(Bucket(cur, t_max, u_max, padded_cost(len(cur), t_max, u_max, vocab, kind)))
    return buckets


def deal_buckets(buckets: Sequence[Bucket], world_size: int) -> List[List[int]]:
    """Greedy cost balancing: heaviest bucket first, each to the currently lightest rank.
    Returns, per rank, the list of bucket indices it owns (deterministic on every rank)."""
    loads = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for b in sorted(range(len(buckets)), key=lambda k: (-buckets[k].cost, k)):
        r = min(range(world_size), key=lambda k: (loads[k], k))
        owned[r].append(b)
        loads[r] += buckets[b].cost
    for o in owned:
        o.sort()
    return owned


def deal_utterances(in_lens: Sequence[int], tgt_lens: Sequence[int], vocab: int, world_size: int,
                    kind: str = "ctc") -> List[List[int]]:
    """Deal UTTERANCES (not buckets) to the ranks, heaviest first to the currently lightest rank, by their own
    (unpadded) logit bytes.  Each rank then cuts its share into length buckets itself (bucket_by_length): with
    thousands of utterances the ranks' loads agree to a fraction of a percent whatever the bucket size is, so the
    buckets can stay large (a launch per bucket fills the GPU) at any number of ranks.  Deterministic."""
    cost = [padded_cost(1, in_lens[i], tgt_lens[i], vocab, kind) for i in range(len(in_lens))]
    loads = [0] * world_size
    owned = [[] for _ in range(world_size)]
    import heapq
    heap = [(0, r) for r in range(world_size)]
    for i in sorted(range(len(cost)), key=lambda k: (-cost[k], k)):
        load, r = heapq.heappop(heap)
        owned[r].append(i)
        heapq.heappush(heap, (load + cost[i], r))
    for o in owned:
        o.sort()
    return owned


def bucketed_pass(kind: str, data, streams, want_grad: bool = True):
    """Loss (+ logit gradient) of every bucket a rank owns, the buckets rotating over `streams` so that a short
    bucket's kernels (one CTA per utterance and sweep direction) run beside its neighbours' instead of leaving SMs
    idle.  data: list of (x, targets, in_len, tgt_len, grad_out) with x (B,T,V) batch-major logits (ctc, star:
    handed to the op as the permuted (T,B,V) view, the way ha/recognizer.py:70 does) or the (B,T,U+1,V) joint
    (rnnt).  Returns [(loss (B,), grad like x or None)] in bucket order; the current stream waits for all of it."""
    from . import ops
    main = torch.cuda.current_stream()
    for st in streams:
        st.wait_stream(main)
    out = []
    for j, (x, tg, il, tl, go) in enumerate(data):
        with torch.cuda.stream(streams[j % len(streams)]):
            if kind == "ctc":
                xv = x.permute(1, 0, 2)
                loss, ws = ops.ctc_fwd(xv, tg, il, tl, True)
                g = ops.ctc_bwd(xv, ws, go, tg.shape[1], True).permute(1, 0, 2) if want_grad else None
            elif kind == "star":
                xv = x.permute(1, 0, 2)
                loss, ws = ops.star_fwd(xv, tg, il, tl, -0.5, True)
                g = ops.star_bwd(xv, ws, go, tg.shape[1], True).permute(1, 0, 2) if want_grad else None
            else:
                loss, ws = ops.rnnt_fwd(x, tg, il, tl, True)
                g = ops.rnnt_bwd(x, ws, go, True) if want_grad else None
            out.append((loss, g))
    for st in streams:
        main.wait_stream(st)
    return out


class LengthBucketBatchSampler(torch.utils.data.Sampler):
    """Drop-in for the reference's DurationBatchSampler (ha/sampler.py:7-29) as the `batch_sampler` of the hac
    training DataLoader (ha/loop.py:502-509): same interface (`data_source.duration(i)`, `max_duration`), same
    growth rule ((len(batch) + 1) * max duration of the batch <= max_duration), but over a LENGTH-SORTED order so the
    padding inside a batch stays small, and sharded: the utterances are first dealt to `world_size` ranks by
    duration (heaviest first to the lightest rank), every rank batches its own share, and the order of the batches
    is reshuffled every epoch from `seed + epoch` (set_epoch, as torch's DistributedSampler).  Every rank yields the
    same NUMBER of batches (the shortest share's count; its leftover batches are dropped like drop_last=True),
    so DDP ranks step in lockstep."""

    def __init__(self, data_source, max_duration=240, rank=0, world_size=1, shuffle=True, seed=0):
        self.data_source = data_source
        self.max_duration = max_duration
        self.rank, self.world_size, self.shuffle, self.seed, self.epoch = rank, world_size, shuffle, seed, 0
        dur = [float(data_source.duration(i)) for i in range(len(data_source))]
        import heapq
        heap = [(0.0, r) for r in range(world_size)]
        shares = [[] for _ in range(world_size)]
        for i in sorted(range(len(dur)), key=lambda k: (-dur[k], k)):
            load, r = heapq.heappop(heap)
            shares[r].append(i)
            heapq.heappush(heap, (load + dur[i], r))
        self._batches = [self._cut(sorted(sh, key=lambda k: (dur[k], k)), dur) for sh in shares]
        self._len = min(len(b) for b in self._batches)

    def _cut(self, order, dur):
        batches, batch, mx = [], [], 0.0
        for i in order:
            new_mx = max(mx, dur[i])
            if batch and (len(batch) + 1) * new_mx > self.max_duration:
                batches.append(batch)
                batch, mx = [i], dur[i]
            else:
                batch.append(i)
                mx = new_mx
        if batch:
            batches.append(batch)
        return batches

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self._len

    def __iter__(self):
        mine = self._batches[self.rank]
        order = list(range(len(mine)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(len(mine), generator=g).tolist()
        for k in order[:self._len]:
            yield list(mine[k])


def shard_batch(n_utts: int, rank: int, world_size: int) -> range:
    """Contiguous split of one already-formed batch: utterances [lo, hi) belong to `rank`."""
    per, rem = divmod(n_utts, world_size)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


def reduce_loss(local_losses: torch.Tensor, weights: torch.Tensor = None, group=None):
    """Global weighted mean of per-utterance losses across ranks with ONE all-reduce of two numbers.

    weights = 1/target_length reproduces ctc_reduce_mean (ha/ctc.py:177-178) over the global batch;
    weights = None is the plain batch mean torchaudio's rnnt_loss 'mean' uses (ha/recognizer.py:125).
    Works on any backend (nccl on GPUs, gloo in the CPU tests); without an initialised process group
    it is the local mean.  Differentiable w.r.t. local_losses.
    """
    import torch.distributed as dist
    w = torch.ones_like(local_losses) if weights is None else weights.to(local_losses.dtype)
    num = (local_losses * w).sum() if weights is not None else local_losses.sum()
    cnt = torch.tensor(float(local_losses.numel()), device=local_losses.device, dtype=local_losses.dtype)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return num / cnt
    buf = torch.stack([num.detach(), cnt])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    # value = global mean; gradient flows only through this rank's own numerator
    return (num - num.detach() + buf[0]) / buf[1]

"""Multi-GPU wiring for the hac training loop (ha/loop.py) — SURVEY §8f rank 3.

The reference trains one process on one GPU: `System` (ha/loop.py:46-142) owns `encoder` and `recognizer` modules,
`train_one_epoch` (ha/loop.py:144-216) calls `self.forward`, scales the loss and steps the optimizer, and `main`
builds a plain shuffling DataLoader (ha/loop.py:502-509).  Nothing of that file is edited; these two helpers give a
maintainer the one-process-per-GPU version of it:

    torch.distributed.init_process_group("nccl")
    haloop_b200.patch_haloop()                                   # CUDA losses behind the reference's call sites
    system = ...                                                 # as ha/loop.py:493 builds it
    haloop_b200.loop.distribute(system, local_rank)              # DDP around encoder and recognizer
    train_loader = haloop_b200.loop.sharded_loader(dataset, Collator(vocab), max_duration=240, num_workers=...)
    for epoch in ...:
        train_loader.batch_sampler.set_epoch(epoch)
        system.train_one_epoch(epoch, global_step, train_loader, valid_loader)

The alignment losses need no collective of their own (utterances are independent; the batch mean inside
`recognizer.forward` is per rank and DDP averages the parameter gradients), so gradient all-reduce is the only
communication, and `sharding.reduce_loss` gives the global mean loss for logging.
"""
import torch
import torch.distributed as dist

from .sharding import LengthBucketBatchSampler


def distribute(system, local_rank=None, **ddp_kwargs):
    """Wrap `system.encoder` and `system.recognizer` (ha/loop.py:57-76) in DistributedDataParallel, in place.
    `System.forward` calls both through `__call__` (ha/loop.py:125-134), which is what DDP hooks; state dicts keep
    working because `make_state_dict` / `load_state_dict` (ha/loop.py:92-111) are given the unwrapped modules back
    through the `module` attribute (see `state_modules`).  CPU modules (gloo) are wrapped without device ids."""
    if not dist.is_initialized():
        raise RuntimeError("distribute() needs torch.distributed.init_process_group() first")
    for name in ("encoder", "recognizer"):
        mod = getattr(system, name)
        if isinstance(mod, torch.nn.parallel.DistributedDataParallel):
            continue
        if not any(p.requires_grad for p in mod.parameters()):
            continue                                                 # frozen module: nothing to all-reduce
        on_cuda = next(mod.parameters()).is_cuda
        kw = dict(ddp_kwargs)
        if on_cuda:
            dev = local_rank if local_rank is not None else torch.cuda.current_device()
            kw.setdefault("device_ids", [dev])
        setattr(system, name, torch.nn.parallel.DistributedDataParallel(mod, **kw))
    return system


def state_modules(system):
    """The unwrapped encoder / recognizer, for checkpoints that must load in a single-process run."""
    unwrap = lambda m: m.module if isinstance(m, torch.nn.parallel.DistributedDataParallel) else m
    return unwrap(system.encoder), unwrap(system.recognizer)


def sharded_loader(dataset, collate_fn, max_duration=240, num_workers=0, shuffle=True, seed=0, rank=None, world_size=None):
    """The training DataLoader of ha/loop.py:502-509 with the length-bucketed, rank-sharded batch sampler
    (`sharding.LengthBucketBatchSampler`: the DurationBatchSampler rule of ha/sampler.py:13-29 over a length-sorted
    order) in place of `batch_size` / `shuffle`.  `dataset.duration(i)` as the reference's datasets provide it."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    sampler = LengthBucketBatchSampler(dataset, max_duration, rank=rank, world_size=world_size, shuffle=shuffle, seed=seed)
    return torch.utils.data.DataLoader(dataset, collate_fn=collate_fn, batch_sampler=sampler, num_workers=num_workers)

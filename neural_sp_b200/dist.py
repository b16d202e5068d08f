"""Data-parallel plumbing for the hot path: utterance sharding and ONE flat gradient all-reduce per step.

Mirrors the reference's DDP semantics: the sampler strides the length-bucketed batch
``indices[rank::world]`` (datasets/asr/sampler.py:61,96); DDP averages gradients and the loss is pre-multiplied by
``num_replicas`` (bin/asr/train.py:423-424), i.e. the effective update is the SUM over ranks of per-rank mean losses.
torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests) is plumbing only."""
import torch
import torch.distributed as dist


def shard_batch(items, rank, world):
    """Rank's slice of a (length-sorted) global batch, strided exactly like the reference sampler."""
    return list(items)[rank::world]


def flat_allreduce_grads(params, op="sum"):
    """All-reduce every parameter gradient with a single collective over one flat buffer; writes the result back."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if op == "mean":
            flat /= dist.get_world_size()
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
    return flat

"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded with no data-path collective;
the only exchange is the PPO gradient all-reduce for the two (PyTorch) policy networks, inserted
between backward() and clip_grad_norm_ (agents/solver.py:195-199, agents/architect.py:138-141)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/MASTER_*). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Env indices [lo, hi) owned by `rank`: contiguous blocks, remainder spread over the first ranks."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class GradBucket:
    """Persistent flat gradient bucket: every parameter's .grad is a VIEW into one contiguous buffer, so the
    all-reduce between backward() and clip_grad_norm_ (agents/solver.py:195-199, agents/architect.py:138-141) is one
    collective on memory the backward pass already wrote -- no per-step torch.cat, no copy-back.

        bucket = GradBucket(net.parameters())
        loss.backward()             # accumulates into the views (zero them with bucket.zero(), not set_to_none)
        work = bucket.allreduce_async()   # NCCL: ReduceOp.AVG on its own stream
        ...                               # independent work (next minibatch's state expansion) overlaps
        bucket.wait(work)
        clip_grad_norm_(...); optimizer.step(); bucket.zero()
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=dt, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.nbytes = self.flat.numel() * self.flat.element_size()

    def zero(self):
        self.flat.zero_()

    def allreduce_async(self):
        """Start the all-reduce (mean over ranks); returns a handle for wait(), or None for a single process."""
        world = _world(self.group)
        if world == 1:
            return None
        avg = dist.get_backend(self.group) == "nccl"
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group, async_op=True)
        return (work, None if avg else world)

    def wait(self, handle):
        if handle is None:
            return
        work, div = handle
        work.wait()
        if div:
            self.flat /= div

    def allreduce(self):
        self.wait(self.allreduce_async())


def allreduce_gradients(params, group=None, average=True):
    """One-shot variant for callers without a GradBucket: flattens, all-reduces, copies back.  (The measured
    loop uses GradBucket; this stays for ad-hoc modules.)"""
    if _world(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def allreduce_sum(t, group=None):
    """Tiny counters (steps, outcomes) summed over ranks; identity for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    return t


def allreduce_max(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t

"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in CPU tests).

The reference is single-device (no tf.distribute / NCCL call sites), so this
is new: the path shards by *batch* — samples are independent through every layer of the
path — and the only exchange step of a training pass is the sum of the weight gradients
(plus BatchNorm statistics when use_bn=True), done here as ONE flat all-reduce per step:
the whole gradient of a HealpyGCNN is 1e2..1e5 floats, i.e. latency-bound, so bucketing by
size would only add launches.
"""

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment; returns
    (rank, world_size, local_rank).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n_items, rank, world):
    """Contiguous [begin, end) share of n_items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_gradients(params, group=None, average=True):
    """Sum (or average) the .grad of `params` over all ranks with a single flat all-reduce.
    Parameters without a gradient contribute zeros so every rank issues the same collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off : off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel()


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def allreduce_max(value, device):
    """Max over ranks of a python float (used for max-over-ranks timing)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class _AllReduceSum(torch.autograd.Function):
    """Differentiable sum over ranks: d(sum_r x_r)/dx_r = 1 on every rank, so the backward pass is the same
    all-reduce applied to the incoming gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _AllReduceSumReplicated(torch.autograd.Function):
    """Sum over ranks whose RESULT is consumed by replicated computation (every rank evaluates the same loss on it):
    the incoming gradient is already the full dL/dy on every rank, so the backward is the identity - summing it over
    the ranks would count the loss once per rank."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def allreduce_sum_replicated(x, group=None):
    """Sum of `x` over all ranks for a replicated consumer (see _AllReduceSumReplicated)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSumReplicated.apply(x, group)


def allreduce_sum_autograd(x, group=None):
    """Sum of `x` over all ranks, differentiable (identity for a single process)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


def batch_statistics(x, dims, group=None, sync=True):
    """Per-channel mean and (biased) variance of `x` over `dims` AND over all ranks: with the batch
    sharded over GPUs, BatchNormalization (gnn_layers.py:53, 152-153) must see the statistics of the global batch
    to reproduce the single-device reference.  One all-reduce of [count, sum, sum of squares] (2F + 1 floats)."""
    # `sync=False`: local statistics even under an initialised process group (a replicated model that the ranks do not
    # run in lock-step would otherwise dead-lock in the all-reduce)
    if not sync or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x.mean(dim=dims), x.var(dim=dims, unbiased=False)
    n = 1
    for d in dims:
        n *= x.shape[d]
    stats = torch.cat([x.sum(dim=dims), (x * x).sum(dim=dims), x.new_full((1,), float(n))])
    stats = allreduce_sum_autograd(stats, group)
    F = x.shape[-1]
    total = stats[2 * F]
    mean = stats[:F] / total
    var = (stats[F : 2 * F] / total - mean * mean).clamp_min(0.0)
    return mean, var

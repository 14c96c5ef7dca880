"""Multi-GPU use of the path: one process per GPU, the batch axis sharded, no data-path collective.

Every voxel is independent given its sample's echo times, so a batch of slices is cut into contiguous
per-rank shards; forward-only operators need no communication at all, and the physics objective needs exactly one
all-reduce of its scalar (NCCL over NVLink on GPUs, gloo in the CPU tests).  The per-rank kernels are told the size
of the GLOBAL batch (inv_n), so the local losses and gradients are already correctly normalised: the sum over ranks of
the local losses is the global mean, and each rank's gradient maps are the global objective's gradient for its shard.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced [start, stop) of `n` items for `rank`; the first n % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard(tensor, rank=None, world=None):
    """This rank's slice of a tensor along the batch axis."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = shard_bounds(tensor.shape[0], rank, world)
    return tensor[a:b]


def sharded_physics_loss(loss_fn, acqs_shard, maps_shard, te_shard, global_elements, group=None, **kw):
    """loss_fn(acqs, maps, te, inv_n=..., **kw) -> scalar tensor normalised by the GLOBAL element count (e.g.
    torch_ops.physics_loss_a2a).  Returns the global objective (all-reduced, detached copy for logging) and the local,
    differentiable term whose backward yields this shard's gradient maps."""
    local = loss_fn(acqs_shard, maps_shard, te_shard, inv_n=1.0 / float(global_elements), **kw)
    total = local.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total.reshape(()), local


def gather_batch(tensor_shard, global_nb, group=None):
    """Optional all-gather of per-shard results (e.g. gradient maps) into one tensor on every rank.  Shards may be
    ragged (global_nb % world != 0): they are padded to the largest shard for the collective and trimmed after."""
    world = dist.get_world_size(group)
    sizes = [b - a for a, b in (shard_bounds(global_nb, r, world) for r in range(world))]
    longest = max(sizes)
    padded = tensor_shard.new_zeros((longest,) + tuple(tensor_shard.shape[1:]))
    padded[: tensor_shard.shape[0]] = tensor_shard
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)

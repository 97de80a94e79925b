"""Data parallelism of the hot path: rays are sharded across ranks, the model is replicated, and the
only exchange of a training step is a SUM all-reduce of the flat gradient buffers.  Inference shards contiguous ray
tiles with no collective.  Backend: torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def all_reduce_gradients(buffers, found_inf=None, group=None):
    """In-place SUM all-reduce of the persistent gradient buffers (hash-table sink, flat MLP gradients) and MAX of the
    GradScaler inf flag so that every rank takes the same skip decision."""
    if world_size(group) == 1:
        return
    for b in buffers:
        dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    if found_inf is not None:
        dist.all_reduce(found_inf, op=dist.ReduceOp.MAX, group=group)


def unscale_factor(loss_scale, world):
    """Multiplier applied to the summed gradients inside the fused optimizer: 1 / (loss_scale * world)."""
    return 1.0 / (float(loss_scale) * int(world))


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n rays for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def interleaved_tiles(n, rank, world, tile=4096):
    """Ray ids of `rank` when the n rays of a frame are dealt out in tiles of `tile` consecutive rays, round robin: every rank
    gets the same mix of image regions (rays that hit the scene and rays that miss it), unlike contiguous N / world slices whose
    edge ranks see mostly background.  Returns an int64 tensor (CPU); concatenating the ranks' results in this order and
    scattering by id restores the frame.  No collective is involved."""
    n, world, tile = int(n), int(world), int(tile)
    ids = torch.arange(n, dtype=torch.int64)
    if world == 1:
        return ids
    return ids[(ids // tile) % world == rank]

"""Ray sharding helpers shared by `_ShardedRender` (single process, several GPUs) and the
one-process-per-GPU launcher in bench.py.  Rays are independent units; shards are the
contiguous `torch.chunk` pieces along the ray axis, so concatenating shard outputs in rank
order reproduces the caller's ray order exactly (what `DataParallel(dim=1)` guarantees in the
reference, src/render/nerf.py:370)."""
import torch


def shard_bounds(n_rays, world):
    """[(start, stop)] per rank with torch.chunk sizes: ceil(n/world) each, last ones may be short/empty."""
    per = -(-n_rays // world) if n_rays > 0 else 0
    out = []
    for r in range(world):
        a = min(r * per, n_rays)
        out.append((a, min(a + per, n_rays)))
    return out


def local_shard(rays, rank, world, dim=1):
    a, b = shard_bounds(rays.shape[dim], world)[rank]
    return rays.narrow(dim, a, b - a)


def broadcast_state(tensors, dist, src=0):
    """One broadcast per tensor of the read-only render state (latent, cameras, weights)."""
    for t in tensors:
        dist.broadcast(t, src=src)


def gather_rays(local, n_rays, dist, rank, world, dst=0, dim=1):
    """Gather per-rank outputs (ragged along `dim`) to `dst` in rank order; returns the full tensor on dst, None elsewhere.
    Shards are padded to the common chunk size so a single fixed-size gather suffices."""
    bounds = shard_bounds(n_rays, world)
    per = max(b - a for a, b in bounds) if bounds else 0
    pad_shape = list(local.shape)
    pad_shape[dim] = per
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r].narrow(dim, 0, bounds[r][1] - bounds[r][0]) for r in range(world)], dim=dim)

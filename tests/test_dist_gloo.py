"""world_size-2 CPU (gloo) test of the multi-process host logic used by bench.py: state
broadcast from rank 0, ray sharding, ragged gather back in ray order."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pixel-nerf_b200", "src")


def _fake_render(rays, state):
    # any deterministic per-ray function of (rays, broadcast state): order errors show up as mismatches
    return torch.stack((rays[..., :3].sum(-1) * state[0], rays[..., 3:6].sum(-1) + state[1], rays[..., 6] * rays[..., 7]), -1)


def _worker(rank, world, port, n_rays, q):
    sys.path.insert(0, SRC)
    from render.sharding import broadcast_state, gather_rays, local_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    rays = torch.rand(1, n_rays, 8, generator=g)
    state = torch.tensor([2.0, -1.0]) if rank == 0 else torch.zeros(2)  # only rank 0 has the real state
    broadcast_state([state], dist, src=0)
    mine = local_shard(rays, rank, world)
    out = _fake_render(mine, state)
    full = gather_rays(out, n_rays, dist, rank, world, dst=0)
    if rank == 0:
        q.put((full - _fake_render(rays, torch.tensor([2.0, -1.0]))).abs().max().item())
    dist.barrier()
    dist.destroy_process_group()


def _run(n_rays):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_rays) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return err


def test_gloo_two_ranks_even_split():
    assert _run(64) == 0.0


def test_gloo_two_ranks_ragged_split():
    assert _run(37) == 0.0

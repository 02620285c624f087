#!/usr/bin/env python
"""Single-process multi-GPU render through the public API: `renderer.bind_parallel(net, [0..N-1])` (what the reference's
eval scripts call with `--gpu_id "0 1 ..."`, src/render/nerf.py:354-371) on the pnr_mgpu_* driver.

    python scripts/bench_sharded.py --gpus 2 [--workload c2] [--frames 4] [--steps 5]

Prints one JSON line: steady-state rays/s over `--frames` frames per step (rays resident on gpus[0], pixels returned to
gpus[0]), and the encode-to-first-pixel latency: `net.encode()` of a NEW object followed by the first render call, i.e.
trunk + re-layout + projection on gpus[0] + the peer copies that refresh the replicas + the first sharded render
(the reference's DataParallel pays its whole-module broadcast on every call instead)."""
import argparse
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pnr_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--workload", default="c2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step (default: one per GPU)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    synth = bench.synth
    cfg = synth.CONFIGS[a.workload]
    dev0 = torch.device("cuda:0")
    torch.cuda.set_device(dev0)
    net, renderer = bench.build_scene(cfg, dev0, "auto")
    gpus = list(range(a.gpus))
    render_par = renderer.bind_parallel(net, gpus, simple_output=True).eval()
    frames = a.frames or a.gpus
    n_rays = frames * bench.WORKLOADS[a.workload]["frame_rays"]
    rays = synth.make_rays(cfg, n_rays, n_target=max(8, frames + 1)).to(dev0)[None]

    def sync():
        for g in gpus:
            torch.cuda.synchronize(g)

    def step():
        with torch.no_grad():
            return render_par(rays)

    for _ in range(a.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    sync()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)

    # encode-to-first-pixel: a new object (new images), then the first sharded frame
    src, _, focal, c = synth.make_cameras(cfg)
    lat = []
    one = rays[:, :bench.WORKLOADS[a.workload]["frame_rays"]]
    for seed in (1, 2, 3):
        images = synth.make_images(cfg, seed=seed)[None].to(dev0)
        sync()
        t0 = time.perf_counter()
        with torch.no_grad():
            net.encode(images, src[None].to(dev0), focal.to(dev0), c=c[None].to(dev0))
            rgb, _ = render_par(one)
        sync()
        lat.append((time.perf_counter() - t0) * 1e3)
    with torch.no_grad():
        sync()
        t0 = time.perf_counter()
        render_par(one)
        sync()
        frame_ms = (time.perf_counter() - t0) * 1e3
    reps = {g: r.refreshes for g, r in getattr(render_par, "_replicas", {}).items()}
    print(json.dumps({"metric": bench.WORKLOADS[a.workload]["metric"], "api": "bind_parallel(net, gpus) single process",
                      "n_gpus": a.gpus, "value": n_rays * a.steps / (ms / 1e3), "unit": "rays/s", "rays_per_step": n_rays,
                      "ms_per_step": ms / a.steps, "wall_ms_per_step": wall / a.steps * 1e3,
                      "encode_to_first_pixel_ms": sorted(lat)[1], "steady_one_frame_ms": frame_ms,
                      "replica_refreshes": reps, "finite": bool(torch.isfinite(rgb).all())}))


if __name__ == "__main__":
    main()

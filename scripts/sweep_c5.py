#!/usr/bin/env python
"""BASELINE.json config C5: ray-batch sweep 4k-1M rays x {1,3,6} source views at C2 geometry (128x128 target,
64+32 samples, d=512), one GPU, resident rays, synthetic latents of the C2 size (64x64).  Prints rays/s and the
algorithmic TFLOP/s per point of the sweep.  Usage (on the GPU box): python scripts/sweep_c5.py > gpurun_out/sweep.txt"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-nerf_b200", "src"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util  # noqa: E402
import golden_util as gu  # noqa: E402
from model import make_model  # noqa: E402
from render import NeRFRenderer  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(gu.synth.CONFIGS["c2"])
net = make_model(gpu_util.model_conf(512))
net.mlp_coarse.load_state_dict(gu.synth.make_mlp_weights(11, 512))
net.mlp_fine.load_state_dict(gu.synth.make_mlp_weights(12, 512))
net = net.to(dev).eval()
renderer = NeRFRenderer(n_coarse=64, n_fine=32, n_fine_depth=16, white_bkgd=True).eval()
print(f"{'NS':>3} {'rays':>8} {'ms':>9} {'rays/s':>10} {'alg TFLOP/s':>12}")
for NS in (1, 3, 6):
    cfg["NS"] = NS
    poses = torch.stack([gu.synth.pose_spherical(40.0 * v, -30.0, 1.3) for v in range(NS)])[None].to(dev)
    net.set_scene(gu.synth.make_latent(3, NS, 64, 64).to(dev), poses, torch.tensor(131.25, device=dev), None, 128, 128)
    par = renderer.bind_parallel(net, [0], simple_output=True).eval()
    fl = gu.synth.flops_per_ray(64, 32, NS, 512)
    for n in (4096, 16384, 65536, 262144, 1048576):
        rays = gu.synth.make_rays(cfg, n, n_target=max(8, n // 16384 + 1)).to(dev)[None]
        with torch.no_grad():
            par(rays[:, :4096])
            torch.cuda.synchronize()
            reps = 3 if n <= 65536 else 1
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                par(rays)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{NS:>3} {n:>8} {ms:>9.2f} {n / ms * 1e3:>10.0f} {n / ms * 1e3 * fl / 1e12:>12.1f}", flush=True)

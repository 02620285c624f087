"""SURVEY 8f-3 rows on the GPU: pnr_gen_rays against the reference-generated fixture and the oracle, pnr_frames_u8
bit-exact against numpy's `(x * 255).astype(uint8)`, and render_frames == the reference's caller loop
(eval/gen_video.py:166-222, :236) byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pixel-nerf_b200")
sys.path.insert(0, os.path.join(PKG, "src"))
pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RAY_TOL = 1e-6   # unit-scale fp32 values; the three-term dot products may round differently from torch's matmul


def test_gen_rays_matches_reference_fixture():
    import pnr_native as pn
    z = np.load(gu.GOLD + "/util_rays.npz")
    poses = torch.from_numpy(z["poses"]).to(DEV)
    rays = pn.gen_rays(poses, 12, 9, 13.5, 13.5, 6.0, 4.5, 0.8, 1.8).view(3, 9, 12, 8).cpu().numpy()
    assert np.abs(rays - z["rays"]).max() < RAY_TOL
    assert np.array_equal(rays[..., :3], z["rays"][..., :3]) and np.array_equal(rays[..., 6:], z["rays"][..., 6:])
    rays_c = pn.gen_rays(poses[:1], 12, 9, 13.5, 14.0, 6.5, 4.0, 0.1, 5.0).view(1, 9, 12, 8).cpu().numpy()
    assert np.abs(rays_c - z["rays_c"]).max() < RAY_TOL


def test_gen_rays_ranges_and_util_entry():
    import pnr_native as pn
    import util
    poses = torch.stack([gu.synth.pose_spherical(a, p, 1.7) for a, p in ((10, -20), (200, -35), (77.7, 5), (-90, -60))])
    W, H, f, c = 37, 23, torch.tensor([41.0, 39.5]), torch.tensor([17.25, 12.5])
    want = gu.oracle.gen_rays(poses, W, H, 41.0, 39.5, 17.25, 12.5, 0.3, 4.0)
    full = util.gen_rays(poses.to(DEV), W, H, f, 0.3, 4.0, c=c)          # the reference's signature, CUDA poses
    assert full.shape == (4, H, W, 8) and full.is_cuda
    assert (full.cpu() - want).abs().max() < RAY_TOL
    flat = want.view(-1, 8)
    for first, count in ((0, 1), (5, 31), (33, 64), (100, 1000), (4 * W * H - 7, 7), (17, 0)):
        part = pn.gen_rays(poses.to(DEV), W, H, 41.0, 39.5, 17.25, 12.5, 0.3, 4.0, first, count)
        assert part.shape == (count, 8)
        if count:
            assert torch.equal(part, full.view(-1, 8)[first:first + count])      # same kernel arithmetic at any offset
            assert (part.cpu() - flat[first:first + count]).abs().max() < RAY_TOL
    with pytest.raises(RuntimeError):
        pn.gen_rays(poses.to(DEV), W, H, 41.0, 39.5, 17.25, 12.5, 0.3, 4.0, 4 * W * H - 3, 8)   # past the grid


def test_gen_rays_large_vs_oracle():
    import pnr_native as pn
    poses = torch.stack([gu.synth.pose_spherical(a, -25.0, 2.5) for a in np.linspace(-180, 180, 6)[:-1]])
    rays = pn.gen_rays(poses.to(DEV), 400, 300, 360.0, 355.0, 200.0, 150.0, 0.1, 5.0).view(5, 300, 400, 8)
    want = gu.oracle.gen_rays(poses, 400, 300, 360.0, 355.0, 200.0, 150.0, 0.1, 5.0)
    assert (rays.cpu() - want).abs().max() < RAY_TOL
    n = rays[..., 3:6].norm(dim=-1)
    assert (n - 1).abs().max() < 1e-6


def test_frames_u8_fixture_and_random():
    import pnr_native as pn
    z = np.load(gu.GOLD + "/frames_u8.npz")
    got = pn.frames_u8(torch.from_numpy(z["rgb"]).to(DEV)).cpu().numpy()
    assert got.dtype == np.uint8 and np.array_equal(got, z["u8"])
    g = torch.Generator().manual_seed(3)
    for n in (1, 2, 3, 5, 4096, 1_000_003):
        x = torch.rand(n, generator=g)
        x[::7] = torch.round(x[::7] * 255) / 255                      # exact k/255 boundaries
        assert np.array_equal(pn.frames_u8(x.to(DEV)).cpu().numpy(), gu.oracle.frames_u8(x))
    assert pn.frames_u8(torch.empty(0, 3, device=DEV)).shape == (0, 3)


def test_render_frames_equals_the_reference_caller_loop():
    import util
    from model import make_model
    from render import NeRFRenderer, render_frames
    conf = util.hocon.parse_file(os.path.join(PKG, "conf", "exp", "srn.conf"))
    conf.put("model.encoder.pretrained", False)
    net = make_model(conf["model"]).to(device=DEV)
    with torch.no_grad():
        for mlp in (net.mlp_coarse, net.mlp_fine):
            for blk in mlp.blocks:
                blk.fc_1.weight.normal_(0, 0.03)
            mlp.lin_out.weight.mul_(0.2)          # keep sigma = relu(2.5 + small) clearly positive: a visible object
            mlp.lin_out.bias[3] = 2.5
    bs = 1000
    renderer = NeRFRenderer.from_conf(conf["renderer"], lindisp=False, eval_batch_size=bs).to(device=DEV)
    render_par = renderer.bind_parallel(net, [0], simple_output=True).eval()
    W, H, NV, z_near, z_far = 40, 30, 2, 0.8, 1.8
    focal = torch.tensor(41.0, device=DEV)
    images = torch.rand(2, 3, H, W) * 2 - 1
    src = torch.stack([util.pose_spherical(a, -30.0, 1.3) for a in (0.0, 40.0)])
    poses = torch.stack([util.pose_spherical(a, -10.0, 1.3) for a in np.linspace(-180, 180, NV + 1)[:-1]], 0).to(DEV)
    with torch.no_grad():
        net.encode(images.unsqueeze(0).to(DEV), src.unsqueeze(0).to(DEV), focal)
        # the reference's loop (gen_video.py:166-222, 236)
        torch.manual_seed(11)
        render_rays = util.gen_rays(poses, W, H, focal, z_near, z_far)
        all_rgb = []
        for rays in torch.split(render_rays.view(-1, 8), bs, dim=0):
            rgb, _depth = render_par(rays[None])
            all_rgb.append(rgb[0])
        frames = torch.cat(all_rgb).view(-1, H, W, 3)
        want = (frames.cpu().numpy() * 255).astype(np.uint8)
        torch.manual_seed(11)
        got = render_frames(render_par, poses, W, H, focal, z_near, z_far, ray_batch_size=bs)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (NV, H, W, 3)
    assert np.array_equal(got.cpu().numpy(), want)
    assert want.min() < 250 and want.max() > want.min()    # not a blank (all-background) or constant frame

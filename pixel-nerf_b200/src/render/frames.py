"""Novel-view frames straight from camera poses: the caller-side loop of the reference's video script
(eval/gen_video.py:166-222 ray generation + split loop + cat, :236 uint8 conversion) with the rays of each batch
generated on the GPU (`pnr_gen_rays`) and the rendered colours written as uint8 frame bytes (`pnr_frames_u8`).

Only the poses (64 B per view) go to the device and 3 B per ray come back, instead of 32 B per ray in and 16 B out.
"""
import torch

import pnr_native
from util.util import _intrinsics


def render_frames(render_par, poses, width, height, focal, z_near, z_far, c=None, ray_batch_size=50000, out=None):
    """poses (NV,4,4) camera-to-world on the render device -> uint8 frames (NV,H,W,3) on that device.

    `render_par` is what `NeRFRenderer.bind_parallel(net, gpus, simple_output=True)` returns (nerf.py:354-371);
    it is called exactly like the reference does, `rgb, depth = render_par(rays[None])` per batch of
    `ray_batch_size` rays (gen_video.py:209-212), in the same pixel order, so a seeded run draws the same noise.
    """
    if not poses.is_cuda:
        raise RuntimeError("render_frames: poses must be on the CUDA device that renders (no CPU fallback)")
    nv = poses.shape[0]
    fx, fy, cx, cy = _intrinsics(width, height, torch.as_tensor(focal).squeeze(), c)
    total = nv * width * height
    if out is None:
        out = torch.empty(nv, height, width, 3, device=poses.device, dtype=torch.uint8)
    elif tuple(out.shape) != (nv, height, width, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise RuntimeError("render_frames: out must be a contiguous uint8 (NV,H,W,3) tensor")
    flat = out.view(-1)
    rays = torch.empty(min(ray_batch_size, max(total, 1)), 8, device=poses.device, dtype=torch.float32)
    poses32 = poses.to(torch.float32).contiguous()
    with torch.no_grad():
        for first in range(0, total, ray_batch_size):
            count = min(ray_batch_size, total - first)
            batch = rays[:count]
            pnr_native.gen_rays(poses32, width, height, fx, fy, cx, cy, z_near, z_far, first, count, out=batch)
            rgb, _depth = render_par(batch[None])
            # first*3 is a multiple of 4 bytes whenever ray_batch_size is a multiple of 4 (the default is)
            dst = flat[first * 3:(first + count) * 3]
            if dst.data_ptr() % 4 == 0:
                pnr_native.frames_u8(rgb[0], out=dst)
            else:
                dst.copy_(pnr_native.frames_u8(rgb[0]).view(-1))
    return out

"""
Synthetic scenes, rays and ResnetFC weights for benchmarks, parity tests and golden
fixtures.  Stand-alone on purpose: pure torch/numpy, no
package-relative imports, so that `oracle/make_golden.py` can load it by path next to
the reference's own `util` package, and `bench.py` / tests can load it next to ours.

Nothing here is on the hot path; it only manufactures inputs of the shapes the
reference's datasets produce:
  * rays       -- layout of `util.gen_rays` (/root/reference/src/util/util.py:238-276):
                  [origin(3), unit dir(3), near, far] per pixel, row-major (NV, H, W).
  * poses      -- `util.pose_spherical` orbit (/root/reference/src/util/util.py:309-324),
                  source views as in eval/gen_video.py:160, targets as in :166-172.
  * weights    -- ResnetFC state_dict keys (/root/reference/src/model/resnetfc.py:88-121)
                  with every `fc_1.weight` re-randomised (the reference zero-inits it,
                  resnetfc.py:39, which would make half the GEMMs vanish) and a positive
                  sigma bias so a useful fraction of samples is opaque.
"""
import math

import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# Named workloads (BASELINE.json "configs"; constants from SURVEY.md section 8d)
# ----------------------------------------------------------------------------------------
CONFIGS = {
    # C1: plumbing case, CPU-runnable
    "c1": dict(W=64, H=64, focal=65.6, z_near=0.8, z_far=1.8, NS=1, n_coarse=64, n_fine=0,
               n_fine_depth=0, d_hidden=128, white_bkgd=True, use_first_pool=True),
    # C2: SRN-car 128x128, 2 source views, 64+32 samples (16 of the 32 are depth samples)
    "c2": dict(W=128, H=128, focal=131.25, z_near=0.8, z_far=1.8, NS=2, n_coarse=64, n_fine=32,
               n_fine_depth=16, d_hidden=512, white_bkgd=True, use_first_pool=True),
    # C3: ShapeNet-NMR 64x64, 1 source view, 64+16 (8 depth samples)
    "c3": dict(W=64, H=64, focal=70.0, z_near=1.2, z_far=4.0, NS=1, n_coarse=64, n_fine=16,
               n_fine_depth=8, d_hidden=512, white_bkgd=True, use_first_pool=False),
    # C4: DTU 400x300, 3 source views, 96+48
    "c4": dict(W=400, H=300, focal=360.0, z_near=0.1, z_far=5.0, NS=3, n_coarse=96, n_fine=48,
               n_fine_depth=16, d_hidden=512, white_bkgd=False, use_first_pool=True),
}

D_LATENT = 512  # resnet34, num_layers=4: 64+64+128+256 (encoder.py:66)
D_IN = 42       # 3 + 6*2*3 positional code + 3 view dirs (models.py:48-60, code.py:17-20)


def flops_per_ray(n_coarse, n_fine, NS, d_hidden, d_in=D_IN, d_latent=D_LATENT):
    """Algorithmic FLOPs per ray (2*MAC) of the reference's fp32 model, SURVEY.md 8(a)/(d)."""
    d = d_hidden
    m_pv = d_in * d + 3 * (d_latent * d + 2 * d * d)
    m_p = 2 * 2 * d * d + 4 * d
    pts = n_coarse + ((n_coarse + n_fine) if n_fine > 0 else 0)
    return 2 * pts * (NS * m_pv + m_p)


# ----------------------------------------------------------------------------------------
# Cameras and rays
# ----------------------------------------------------------------------------------------
def pose_spherical(theta_deg, phi_deg, radius):
    """Camera-to-world matrix of an orbit camera; same convention as the reference's
    util.pose_spherical (translate along z, tilt by phi, orbit by theta, axis flip)."""
    th = theta_deg / 180.0 * np.pi
    ph = phi_deg / 180.0 * np.pi
    t = torch.eye(4, dtype=torch.float32)
    t[2, 3] = radius
    rp = torch.tensor([[1, 0, 0, 0],
                       [0, np.cos(ph), -np.sin(ph), 0],
                       [0, np.sin(ph), np.cos(ph), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[np.cos(th), 0, -np.sin(th), 0],
                       [0, 1, 0, 0],
                       [np.sin(th), 0, np.cos(th), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]],
                        dtype=torch.float32)
    return flip @ (rt @ (rp @ t))


def gen_rays(poses, width, height, focal, z_near, z_far, c=None):
    """(NV,4,4) c2w poses -> (NV,H,W,8) rays.  Pixel (y,x) looks along
    normalise([(x-cx)/fx, -(y-cy)/fy, -1]) rotated into the world."""
    nv = poses.shape[0]
    dev = poses.device
    f = torch.as_tensor(focal, dtype=torch.float32).reshape(-1)
    fx, fy = (float(f[0]), float(f[0])) if f.numel() == 1 else (float(f[0]), float(f[1]))
    if c is None:
        cx, cy = width * 0.5, height * 0.5
    else:
        cc = torch.as_tensor(c, dtype=torch.float32).reshape(-1)
        cx, cy = float(cc[0]), float(cc[1])
    ys = (torch.arange(height, dtype=torch.float32) - cy).to(dev) / fy
    xs = (torch.arange(width, dtype=torch.float32) - cx).to(dev) / fx
    Y = ys[:, None].expand(height, width)
    X = xs[None, :].expand(height, width)
    d = torch.stack((X, -Y, -torch.ones_like(X)), dim=-1)
    d = d / torch.norm(d, dim=-1).unsqueeze(-1)
    dirs = torch.matmul(poses[:, None, None, :3, :3],
                        d[None].expand(nv, -1, -1, -1).unsqueeze(-1))[..., 0]
    origins = poses[:, None, None, :3, 3].expand(-1, height, width, -1)
    near = torch.full((nv, height, width, 1), float(z_near), device=dev)
    far = torch.full((nv, height, width, 1), float(z_far), device=dev)
    return torch.cat((origins, dirs, near, far), dim=-1)


def make_cameras(cfg, n_target=8):
    """Source poses (NS,4,4), target poses (n_target,4,4), focal (scalar tensor), c (2,)."""
    r = (cfg["z_near"] + cfg["z_far"]) * 0.5
    src = torch.stack([pose_spherical(40.0 * i, -30.0, r) for i in range(cfg["NS"])])
    angles = np.linspace(-180, 180, n_target + 1)[:-1]
    tgt = torch.stack([pose_spherical(float(a), -10.0, r) for a in angles])
    focal = torch.tensor(cfg["focal"], dtype=torch.float32)
    c = torch.tensor([cfg["W"] * 0.5, cfg["H"] * 0.5], dtype=torch.float32)
    return src, tgt, focal, c


def make_images(cfg, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(cfg["NS"], 3, cfg["H"], cfg["W"], generator=g) * 2.0 - 1.0


def make_rays(cfg, n_rays, n_target=8):
    """First n_rays rays (tiled if needed) of the target orbit, flattened (n_rays, 8)."""
    _, tgt, focal, c = make_cameras(cfg, n_target)
    per = cfg["W"] * cfg["H"]
    need = min(n_target, (n_rays + per - 1) // per)
    rays = gen_rays(tgt[:need], cfg["W"], cfg["H"], focal, cfg["z_near"], cfg["z_far"], c)
    rays = rays.reshape(-1, 8)
    if rays.shape[0] < n_rays:
        reps = (n_rays + rays.shape[0] - 1) // rays.shape[0]
        rays = rays.repeat(reps, 1)
    return rays[:n_rays].contiguous()


def make_latent(seed, NS, Hl, Wl, d_latent=D_LATENT):
    """Stand-in for the encoder's pre-upsampled pyramid (encoder.py:150-160) with the
    statistics the random-init resnet34 produces at C2 (mean 10.9, std 18.9):
    non-negative, heavy-tailed.  Used where running the conv trunk is beside the point."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(NS, d_latent, Hl, Wl, generator=g) * 19.0 + 4.0
    return torch.clamp_min(x, 0.0).contiguous()


# ----------------------------------------------------------------------------------------
# ResnetFC weights
# ----------------------------------------------------------------------------------------
def make_mlp_weights(seed, d_hidden, d_in=D_IN, d_latent=D_LATENT, n_blocks=5, combine_layer=3,
                     d_out=4, bias_std=0.05, sigma_bias=2.0):
    """state_dict-shaped dict for one ResnetFC (keys as resnetfc.py registers them)."""
    g = torch.Generator().manual_seed(seed)

    def kaiming(out_f, in_f, gain=1.0):
        return torch.randn(out_f, in_f, generator=g) * (gain * math.sqrt(2.0 / in_f))

    def bias(n):
        return torch.randn(n, generator=g) * bias_std

    sd = {}
    sd["lin_in.weight"] = kaiming(d_hidden, d_in)
    sd["lin_in.bias"] = bias(d_hidden)
    sd["lin_out.weight"] = kaiming(d_out, d_hidden)
    b = bias(d_out)
    b[3] = sigma_bias
    sd["lin_out.bias"] = b
    for i in range(n_blocks):
        sd[f"blocks.{i}.fc_0.weight"] = kaiming(d_hidden, d_hidden)
        sd[f"blocks.{i}.fc_0.bias"] = bias(d_hidden)
        sd[f"blocks.{i}.fc_1.weight"] = kaiming(d_hidden, d_hidden, gain=0.5)
        sd[f"blocks.{i}.fc_1.bias"] = bias(d_hidden)
    for i in range(min(combine_layer, n_blocks)):
        sd[f"lin_z.{i}.weight"] = kaiming(d_hidden, d_latent)
        sd[f"lin_z.{i}.bias"] = bias(d_hidden)
    return sd


BENCH_LATENT_GAIN = 0.01


def bench_mlp_weights(seed, d_hidden, latent_gain=BENCH_LATENT_GAIN):
    """ResnetFC weights for BENCHMARK scenes: `make_mlp_weights` with every `lin_z.*.weight` scaled by `latent_gain`.
    A random-init resnet34 trunk produces a latent of rms ~ 22 with a large positive mean; through
    kaiming-scale lin_z layers that drives every output far into saturation -- with these seeds sigma = relu(< 0) = 0
    at EVERY sample, i.e. a blank white frame, on which a parity check is vacuous.  At 0.01 the latent injection is
    O(0.3) per block, next to the O(1) positional-code path: semi-transparent volumes (mean opacity 0.36 / 0.83 / 0.68
    on C2 / C3 / C4), unsaturated colours, depth that varies over the frame.  The arithmetic (and its cost) is
    unchanged; golden fixtures keep using `make_mlp_weights`."""
    sd = make_mlp_weights(seed, d_hidden)
    for k in sd:
        if k.startswith("lin_z.") and k.endswith(".weight"):
            sd[k] = sd[k] * latent_gain
    return sd


def weights_checksum(sd):
    """Order-independent fingerprint used by golden fixtures to detect RNG drift."""
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return tot


def draw_noise(seed, B, n_coarse, n_fine, n_fine_depth, device="cpu"):
    """The four draws NeRFRenderer.forward makes, in its order (SURVEY A.6;
    nerf.py:111,135,141,158), from one generator."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {"u_coarse": torch.rand(B, n_coarse, generator=g, device=device)}
    kf = n_fine - n_fine_depth
    if n_fine > 0 and kf > 0:
        out["u_fine"] = torch.rand(B, kf, generator=g, device=device)
        out["u_fine_jit"] = torch.rand(B, kf, generator=g, device=device)
    if n_fine > 0 and n_fine_depth > 0:
        out["n_depth"] = torch.randn(B, n_fine_depth, generator=g, device=device)
    return out

"""TEST INFRASTRUCTURE: drive the g++-compiled, host-emulated SIMT translation units (tests/cuda_emu) through the
same C ABI with CPU tensors.  The product path never loads this library."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pixel-nerf_b200", "src"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import pnr_native as pn  # noqa: E402  (struct layouts and signatures only; pn.lib() is never called here)

_emu = None


def lib():
    global _emu
    if _emu is None:
        import build_emu
        # PNR_EMU_LIB: an alternative build of the same sources (e.g. with -fsanitize=address)
        _emu = pn.declare(C.CDLL(os.environ.get("PNR_EMU_LIB") or build_emu.build()))
    return _emu


def ok(rc):
    assert rc == 0, lib().pnr_last_error().decode()


def ptr(t):
    if t is None:
        return None
    assert t.dtype == torch.float32 and t.is_contiguous() and not t.is_cuda
    return C.c_void_p(t.data_ptr())


def mlp_struct(sd, d_hidden, n_blocks=5, combine_layer=3, d_in=42, d_latent=512, d_out=4):
    m = pn.PnrMlp()
    m.d_in, m.d_latent, m.d_hidden, m.d_out = d_in, d_latent, d_hidden, d_out
    m.n_blocks, m.combine_layer = n_blocks, combine_layer
    m.lin_in_w, m.lin_in_b = ptr(sd["lin_in.weight"]), ptr(sd["lin_in.bias"])
    m.lin_out_w, m.lin_out_b = ptr(sd["lin_out.weight"]), ptr(sd["lin_out.bias"])
    for i in range(n_blocks):
        m.fc0_w[i], m.fc0_b[i] = sd[f"blocks.{i}.fc_0.weight"].data_ptr(), sd[f"blocks.{i}.fc_0.bias"].data_ptr()
        m.fc1_w[i], m.fc1_b[i] = sd[f"blocks.{i}.fc_1.weight"].data_ptr(), sd[f"blocks.{i}.fc_1.bias"].data_ptr()
    for i in range(min(combine_layer, n_blocks)):
        m.lin_z_w[i], m.lin_z_b[i] = sd[f"lin_z.{i}.weight"].data_ptr(), sd[f"lin_z.{i}.bias"].data_ptr()
    return m


def scene_struct(case, state, keep):
    """PnrScene over CPU tensors for a golden case; `keep` collects the tensors that must outlive the call."""
    cfg = case["cfg"]
    lat = case["latent"]
    V, Cc, Hl, Wl = lat.shape
    nhwc = lat.permute(0, 2, 3, 1).contiguous()
    poses, focal, c = state["poses"].contiguous(), state["focal"].contiguous(), state["c"].contiguous()
    keep += [nhwc, poses, focal, c]
    s = pn.PnrScene()
    s.latent_nhwc, s.poses, s.focal, s.c = ptr(nhwc), ptr(poses), ptr(focal), ptr(c)
    s.n_focal, s.n_c = focal.shape[0], c.shape[0]
    s.SB, s.NS, s.Hl, s.Wl, s.C = cfg["SB"], cfg["NS"], Hl, Wl, Cc
    s.image_w, s.image_h = float(cfg["W"]), float(cfg["H"])
    s.scale_x = float(Wl) / (float(Wl) - 1.0) * 2.0
    s.scale_y = float(Hl) / (float(Hl) - 1.0) * 2.0
    return s

"""pnr_mgpu_render (csrc/pnr_mgpu.cu, the single-process multi-GPU driver behind `bind_parallel(net, gpus)`) on the host
emulator: the sharding arithmetic that no 1-GPU box can exercise -- torch.chunk bounds incl. ragged and empty shards,
strided staging of rays and outputs for SB > 1, in-place ("peer-stored") final pixels for SB = 1, optional outputs --
checked bit for bit against ONE pnr_render call over all rays with the same per-ray draws.  (Replaces
nn.DataParallel(dim=1), reference src/render/nerf.py:354-371: the gathered ray order must be the caller's.)"""
import ctypes as C

import pytest
import torch

import emu_util as eu
import golden_util as gu

pn = eu.pn


def _cfg(case):
    cfg = case["cfg"]
    rc = pn.PnrRenderCfg()
    rc.n_coarse, rc.n_fine, rc.n_fine_depth, rc.depth_std = cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"], 0.01
    rc.white_bkgd, rc.engine = int(bool(cfg["white_bkgd"])), 1
    return rc


def _noise(nz, Kc, Kf, Kfd, keep):
    lin = torch.linspace(0, 1 - 1.0 / Kc, Kc)
    noise = pn.PnrNoise()
    noise.lin_steps, noise.u_coarse = eu.ptr(lin), eu.ptr(nz["u_coarse"])
    if Kf - Kfd > 0:
        noise.u_fine, noise.u_fine_jit = eu.ptr(nz["u_fine"]), eu.ptr(nz["u_fine_jit"])
    if Kf > 0 and Kfd > 0:
        noise.n_depth = eu.ptr(nz["n_depth"])
    keep += [lin, nz]
    return noise


def _outputs(R, Kc, Kf, want_extras):
    t = dict(rgb_coarse=torch.full((R, 3), -7.0), depth_coarse=torch.full((R,), -7.0))
    if want_extras:
        t.update(weights_coarse=torch.full((R, Kc), -7.0), z_coarse=torch.full((R, Kc), -7.0))
    if Kf > 0:
        t.update(rgb_fine=torch.full((R, 3), -7.0), depth_fine=torch.full((R,), -7.0))
        if want_extras:
            t.update(weights_fine=torch.full((R, Kc + Kf), -7.0), z_fine=torch.full((R, Kc + Kf), -7.0))
    o = pn.PnrRenderOut()
    for k, v in t.items():
        setattr(o, k, eu.ptr(v))
    return o, t


@pytest.mark.parametrize("name,devices,want_extras", [("tiny", [0, 1, 2], False),     # SB = 1: in-place final pixels
                                                       ("tiny", [0, 0], True),          # no peer access: staged copies
                                                       ("sb2_d", [0, 1, 2], True),      # SB = 2: strided staging
                                                       ("tiny_sb2", [0, 1, 2, 3, 4, 5, 6, 7], False)])
def test_sharded_render_equals_one_call(name, devices, want_extras):
    case = gu.load_case(name)
    cfg = case["cfg"]
    SB, B, Kc, Kf, Kfd = cfg["SB"], cfg["B"], cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"]
    R, n = SB * B, len(devices)
    keep = []
    scene = eu.scene_struct(case, gu.oracle_state(case), keep)
    mc = eu.mlp_struct(case["wc"], cfg["d_hidden"])
    mf = eu.mlp_struct(case["wf"], cfg["d_hidden"]) if case["wf"] is not None else None
    rc = _cfg(case)
    L = eu.lib()
    rays = case["rays"].contiguous()
    nz = {k: v.contiguous() for k, v in case["noise"].items()}
    # ---- one call over all rays
    o_ref, t_ref = _outputs(R, Kc, Kf, want_extras)
    nbytes = L.pnr_render_workspace_bytes(scene, mc, mf, rc, B)
    ws = torch.empty(nbytes, dtype=torch.uint8)
    eu.ok(L.pnr_render(scene, mc, mf, rc, eu.ptr(rays), _noise(nz, Kc, Kf, Kfd, keep), o_ref, B, ws.data_ptr(), nbytes, None))
    # ---- the same rays through the multi-GPU driver
    h = C.c_void_p()
    eu.ok(L.pnr_mgpu_create((C.c_int32 * n)(*devices), n, C.byref(h)))
    assert L.pnr_mgpu_size(h) == n
    o0, t0 = _outputs(R, Kc, Kf, want_extras)
    shards = (pn.PnrShard * n)()
    per = -(-B // n)
    for i in range(n):
        a, b = min(B, per * i), min(B, per * (i + 1))
        Bi = b - a
        if Bi <= 0:
            continue                                  # torch.chunk leaves trailing devices without rays
        sub = {k: v.reshape(SB, B, -1)[:, a:b].reshape(SB * Bi, -1).contiguous() for k, v in nz.items()}
        noise = _noise(sub, Kc, Kf, Kfd, keep)
        st_o, st_t = _outputs(SB * Bi, Kc, Kf, want_extras)
        wsb = L.pnr_render_workspace_bytes(scene, mc, mf, rc, Bi)
        wsi = torch.empty(wsb, dtype=torch.uint8)
        stage_rays = torch.full((SB, Bi, 8), float("nan"))
        sh = shards[i]
        sh.scene, sh.mlp_coarse = C.pointer(scene), C.pointer(mc)
        sh.mlp_fine = C.pointer(mf) if mf is not None else None
        sh.noise = C.pointer(noise)
        sh.workspace, sh.workspace_bytes = wsi.data_ptr(), wsb
        sh.rays_stage = eu.ptr(stage_rays)
        sh.stage = st_o
        keep += [noise, st_t, wsi, stage_rays]
    eu.ok(L.pnr_mgpu_render(h, shards, rc, eu.ptr(rays), o0, B, None))
    eu.ok(L.pnr_mgpu_destroy(h))
    for k in t_ref:
        assert torch.equal(t0[k], t_ref[k]), k


def test_mgpu_rejects_incomplete_shards():
    L = eu.lib()
    h = C.c_void_p()
    eu.ok(L.pnr_mgpu_create((C.c_int32 * 2)(0, 1), 2, C.byref(h)))
    shards = (pn.PnrShard * 2)()
    rc = pn.PnrRenderCfg()
    rc.n_coarse = 4
    o = pn.PnrRenderOut()
    dummy = torch.zeros(8)
    assert L.pnr_mgpu_render(h, shards, rc, eu.ptr(dummy), o, 1, None) < 0
    assert b"shard" in L.pnr_last_error()
    assert L.pnr_mgpu_create(None, 2, C.byref(h)) < 0
    eu.ok(L.pnr_mgpu_destroy(h))

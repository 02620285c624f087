"""Pins oracle/pnr_oracle.py against golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, see oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as gu


@pytest.mark.parametrize("name", gu.ORACLE_CASE_NAMES)
def test_state_matches_reference(name):
    case = gu.load_case(name)
    st = gu.oracle_state(case)
    assert torch.equal(st["poses"], case["ref"]["ref_state_poses"])
    assert torch.equal(st["focal"], case["ref"]["ref_state_focal"])
    assert torch.equal(st["c"].reshape(-1), case["ref"]["ref_state_c"].reshape(-1))


@pytest.mark.parametrize("name", gu.ORACLE_CASE_NAMES)
def test_field_matches_reference(name):
    case = gu.load_case(name)
    st = gu.oracle_state(case)
    ref = case["ref"]
    for key, w in (("field_coarse", case["wc"]), ("field_fine", case["wf"] or case["wc"])):
        out = gu.oracle.field_eval(ref["field_xyz"], ref["field_dirs"], st, case["latent"], w,
                                   case["cfg"]["NS"])
        scale = 1.0 + ref[key].abs()
        assert ((out - ref[key]).abs() / scale).max() < 2e-5


@pytest.mark.parametrize("name", gu.ORACLE_CASE_NAMES)
def test_render_matches_reference(name):
    case = gu.load_case(name)
    res = gu.oracle_render(case)
    ref = case["ref"]
    assert torch.equal(res["coarse"]["z"], ref["z_coarse"])  # RNG replay is exact
    assert (res["coarse"]["rgb"] - ref["coarse_rgb"].reshape(-1, 3)).abs().max() < 1e-5
    assert (res["coarse"]["depth"] - ref["coarse_depth"].reshape(-1)).abs().max() < 1e-5
    K = res["coarse"]["weights"].shape[-1]
    assert (res["coarse"]["weights"] - ref["coarse_weights"].reshape(-1, K)).abs().max() < 1e-5
    if case["cfg"]["n_fine"] > 0:
        assert (res["fine"]["z"] - ref["z_fine"]).abs().max() < 1e-5
        assert (res["fine"]["rgb"] - ref["fine_rgb"].reshape(-1, 3)).abs().max() < 1e-5
        assert (res["fine"]["depth"] - ref["fine_depth"].reshape(-1)).abs().max() < 1e-5


def test_gather_restatement_equals_grid_sample():
    """bilinear_border_gather == F.grid_sample(align_corners=True, border) incl. off-image uv."""
    g = torch.Generator().manual_seed(3)
    latent = torch.randn(2, 16, 5, 7, generator=g)
    uv = (torch.rand(2, 200, 2, generator=g) - 0.25) * torch.tensor([20.0, 14.0]) * 1.5
    image_shape = torch.tensor([14.0, 10.0])
    ours = gu.oracle.bilinear_border_gather(latent, uv, image_shape)
    scale = gu.oracle.latent_scaling(latent) / image_shape
    grid = (uv * scale - 1.0).unsqueeze(2)
    ref = torch.nn.functional.grid_sample(latent, grid, align_corners=True, mode="bilinear",
                                          padding_mode="border")[:, :, :, 0].transpose(1, 2)
    assert (ours - ref).abs().max() < 1e-5


def test_util_rays_fixture():
    z = np.load(gu.GOLD + "/util_rays.npz")
    poses = torch.stack([gu.synth.pose_spherical(a, p, 1.3) for a, p in ((0, -30), (40, -30), (123.4, -10))])
    assert np.array_equal(poses.numpy(), z["poses"])
    rays = gu.synth.gen_rays(poses, 12, 9, torch.tensor(13.5), 0.8, 1.8)
    assert np.array_equal(rays.numpy(), z["rays"])  # bit-exact ray layout and order
    rays_c = gu.synth.gen_rays(poses[:1], 12, 9, torch.tensor([13.5, 14.0]), 0.1, 5.0,
                               c=torch.tensor([6.5, 4.0]))
    assert np.array_equal(rays_c.numpy(), z["rays_c"])


def test_oracle_gen_rays_is_the_reference():
    """oracle.gen_rays against rays produced by the reference's util.gen_rays (make_golden.util_fixture)."""
    z = np.load(gu.GOLD + "/util_rays.npz")
    poses = torch.from_numpy(z["poses"])
    assert np.array_equal(gu.oracle.gen_rays(poses, 12, 9, 13.5, 13.5, 6.0, 4.5, 0.8, 1.8).numpy(), z["rays"])
    assert np.array_equal(gu.oracle.gen_rays(poses[:1], 12, 9, 13.5, 14.0, 6.5, 4.0, 0.1, 5.0).numpy(), z["rays_c"])


def test_oracle_frames_u8_fixture():
    z = np.load(gu.GOLD + "/frames_u8.npz")
    assert np.array_equal(gu.oracle.frames_u8(torch.from_numpy(z["rgb"])), z["u8"])
    assert z["u8"].reshape(-1)[:4].tolist() == [0, 255, 255, 127]


def test_sb2_d_is_not_degenerate():
    ref = gu.load_case("sb2_d")["ref"]
    assert ref["coarse_rgb"].std() > 0.1 and ref["fine_rgb"].std() > 0.1 and ref["fine_weights"].max() > 0.5


@pytest.mark.parametrize("name", gu.GRAD_CASE_NAMES)
def test_backward_oracle_matches_reference_gradients(name):
    """SURVEY 8f-1 groundwork: autograd through the oracle's training loss (train/train.py:199-215) reproduces the
    gradients the reference itself computes for every MLP parameter and the latent, incl. the path from the fine
    loss through the depth-centred samples into the coarse MLP (nerf.py:289-291 does not detach the depth)."""
    case, g = gu.load_case(name), gu.load_grad_case(name)
    cfg = case["cfg"]
    lat = case["latent"].clone().requires_grad_(True)
    wc = {k: v.clone().requires_grad_(True) for k, v in case["wc"].items()}
    wf = None if case["wf"] is None else {k: v.clone().requires_grad_(True) for k, v in case["wf"].items()}
    loss = gu.oracle.train_loss(case["rays"], g["rgb_gt"], case["noise"], gu.oracle_state(case), lat, wc, wf,
                                cfg["NS"], cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"],
                                white_bkgd=bool(cfg["white_bkgd"]), eval_batch_size=cfg["eval_batch_size"])
    assert abs(loss.item() - g["loss"]) < 1e-6
    loss.backward()

    def close(a, ref):
        return (a - ref).abs().max() <= 1e-4 * ref.abs().max() + 1e-9

    assert g["g_latent"].abs().max() > 0 and close(lat.grad, g["g_latent"])
    for k, ref in g["gc"].items():
        assert close(wc[k].grad, ref), ("coarse", k)
    for k, ref in g["gf"].items():
        assert close(wf[k].grad, ref), ("fine", k)
    assert g["gc"]["blocks.4.fc_1.weight"].abs().max() > 0 and g["gf"]["lin_in.weight"].abs().max() > 0

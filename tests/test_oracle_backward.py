"""oracle/pnr_backward.py (hand-derived backward of the training loss, no autograd) against autograd through the
oracle and against the gradients the reference produced itself (tests/golden/grad_*.npz).  CPU only."""
import os

import pytest
import torch

import golden_util as gu

bw = gu.load_by_path("pnr_backward", os.path.join(gu.ROOT, "oracle", "pnr_backward.py"))


def rel(a, ref):
    return ((a - ref).abs().max() / (ref.abs().max() + 1e-20)).item()


def manual(case, gt):
    cfg = case["cfg"]
    return bw.train_loss_backward(case["rays"], gt, case["noise"], gu.oracle_state(case), case["latent"], case["wc"],
                                  case["wf"], cfg["NS"], cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"],
                                  white_bkgd=bool(cfg["white_bkgd"]))


@pytest.mark.parametrize("name", ["tiny", "sb2_d", "ns1_coarse_only", "tiny_sb2"])
def test_manual_backward_equals_autograd(name):
    case = gu.load_case(name)
    cfg = case["cfg"]
    g = torch.Generator().manual_seed(9)
    gt = torch.rand(cfg["SB"], cfg["B"], 3, generator=g)
    lat = case["latent"].clone().requires_grad_(True)
    wc = {k: v.clone().requires_grad_(True) for k, v in case["wc"].items()}
    wf = None if case["wf"] is None else {k: v.clone().requires_grad_(True) for k, v in case["wf"].items()}
    loss = gu.oracle.train_loss(case["rays"], gt, case["noise"], gu.oracle_state(case), lat, wc, wf, cfg["NS"],
                                cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"],
                                white_bkgd=bool(cfg["white_bkgd"]), eval_batch_size=cfg["eval_batch_size"])
    loss.backward()
    m_loss, g_c, g_f, d_lat = manual(case, gt)
    assert abs(m_loss.item() - loss.item()) < 1e-6
    if name == "tiny_sb2":          # all-transparent scene: every gradient is exactly zero
        assert all(float(v.abs().max()) == 0.0 for v in g_c.values()) and float(d_lat.abs().max()) == 0.0
        return
    assert rel(d_lat, lat.grad) < 2e-5
    for k, v in wc.items():
        assert rel(g_c[k], v.grad) < 2e-5, ("coarse", k)
    if wf is not None:
        for k, v in wf.items():
            assert rel(g_f[k], v.grad) < 2e-5, ("fine", k)


@pytest.mark.parametrize("name", gu.GRAD_CASE_NAMES)
def test_manual_backward_equals_reference_gradients(name):
    case, ref = gu.load_case(name), gu.load_grad_case(name)
    m_loss, g_c, g_f, d_lat = manual(case, ref["rgb_gt"])
    assert abs(m_loss.item() - ref["loss"]) < 1e-6
    assert rel(d_lat, ref["g_latent"]) < 1e-4
    for k, v in ref["gc"].items():
        assert rel(g_c[k], v) < 1e-4, ("coarse", k)
    for k, v in ref["gf"].items():
        assert rel(g_f[k], v) < 1e-4, ("fine", k)


def test_coarse_mlp_receives_gradient_from_the_fine_loss():
    """nerf.py:289-291: the depth-centred samples are built from the un-detached coarse depth."""
    case = gu.load_case("sb2_d")
    cfg = case["cfg"]
    gt = torch.rand(cfg["SB"], cfg["B"], 3, generator=torch.Generator().manual_seed(2))
    _, g_both, _, _ = bw.train_loss_backward(case["rays"], gt, case["noise"], gu.oracle_state(case), case["latent"],
                                             case["wc"], case["wf"], cfg["NS"], cfg["n_coarse"], cfg["n_fine"],
                                             cfg["n_fine_depth"], white_bkgd=False, lambda_coarse=0.0)
    assert g_both["blocks.4.fc_1.weight"].abs().max() > 0     # lambda_coarse = 0: only the fine loss is left

"""Debug helper: one small tensor-engine field evaluation against the oracle.  Run with PNR_TC_NO_TRAP=1 so that a
barrier timeout records its tag (printed as `status`) instead of trapping the launch."""
import os
os.environ.setdefault("PNR_TC_NO_TRAP", "1")
import sys, os, torch
sys.path.insert(0, "tests"); sys.path.insert(0, "pixel-nerf_b200/src")
import golden_util as gu, gpu_util, pnr_native as pn
case = gu.load_case("c3_small")
cfg = case["cfg"]
net = gpu_util.build_net(case, engine="tc")
g = torch.Generator().manual_seed(40)
P = 40
xyz = (torch.rand(cfg["SB"], P, 3, generator=g) - 0.5) * 2.4
dirs = torch.nn.functional.normalize(torch.randn(cfg["SB"], P, 3, generator=g), dim=-1)
with torch.no_grad():
    out = net(xyz.cuda(), coarse=True, viewdirs=dirs.cuda())
print("status", pn.tc_status())
ref = gu.oracle.field_eval(xyz, dirs, gu.oracle_state(case), case["latent"], case["wc"], cfg["NS"])
out = out.cpu()
print("out[0,:4]", out[0, :4]); print("ref[0,:4]", ref[0, :4])
print("max rgb err", (out[..., :3] - ref[..., :3]).abs().max().item(), "sigma rel", ((out[..., 3]-ref[..., 3]).abs()/(1+ref[...,3].abs())).max().item())

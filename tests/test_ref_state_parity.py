"""State producer and checkpoint compatibility against the UNMODIFIED reference (run in its own interpreter by
tests/ref_probe.py):

  * a19: `SpatialEncoder.forward` / `PixelNeRFNet.encode` -- same state_dict, same images -> the same latent, camera
    state and `index()` values as the reference (src/model/encoder.py:111-164, src/model/models.py:89-144);
  * f-4: a checkpoint written by the reference's own `save_weights` strict-loads here and gives the reference's field
    values; a checkpoint written here strict-loads in the reference (src/model/models.py:268-316).
"""
import os
import subprocess
import sys

import pytest
import torch

import dropin_util as du
import gpu_util

PROBE = os.path.join(du.ROOT, "tests", "ref_probe.py")
needs_ref = pytest.mark.skipif(du.reference_root() is None, reason="no reference checkout (/root/reference or baseline/_ref)")


def probe(*argv):
    env = dict(os.environ)
    env["PIXELNERF_REF"] = du.reference_root()
    r = subprocess.run([sys.executable, PROBE, *argv], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


class Args:
    def __init__(self, d, name, resume=True):
        self.checkpoints_path, self.name, self.resume = d, name, resume


@needs_ref
def test_encoder_and_encode_state_match_the_reference(tmp_path):
    out = str(tmp_path / "enc.pt")
    probe("encoder", out)
    cases = torch.load(out)
    from model import make_model
    for name, ref in cases.items():
        net = make_model(gpu_util.model_conf(64, ref["use_first_pool"])).eval()
        missing = net.load_state_dict(ref["state_dict"], strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        with torch.enable_grad():     # CPU tensors: the composed-torch path (encode() itself has no fused part)
            net.encode(ref["images"], ref["poses"], ref["focal"], c=ref["c"])
            idx = net.encoder.index(ref["uv"], None, net.image_shape)
        lat = net.encoder.latent.detach()
        assert lat.shape == ref["latent"].shape, name
        assert (lat - ref["latent"]).abs().max() <= 1e-6 * ref["latent"].abs().max(), name
        assert torch.equal(net.encoder.latent_scaling, ref["latent_scaling"])
        assert torch.equal(net.poses, ref["poses_state"])
        assert torch.equal(net.focal, ref["focal_state"]) and torch.equal(net.c, ref["c_state"])
        assert torch.equal(net.image_shape, ref["image_shape"])
        assert net.num_views_per_obj == ref["num_views_per_obj"]
        assert (idx.detach() - ref["index"]).abs().max() <= 1e-5 * ref["index"].abs().max()


@needs_ref
def test_checkpoints_round_trip_with_the_reference(tmp_path):
    d = str(tmp_path)
    probe("checkpoint", d)
    assert os.path.exists(os.path.join(d, "probe", "pixel_nerf_latest"))
    assert os.path.exists(os.path.join(d, "probe", "pixel_nerf_backup"))
    io = torch.load(os.path.join(d, "probe_io.pt"))
    from model import make_model
    net = make_model(gpu_util.model_conf(512, True)).eval()
    assert list(net.state_dict().keys()) == io["keys"]            # same names, same order
    assert net.load_weights(Args(d, "probe"), strict=True) is net
    with torch.enable_grad():                                      # CPU: composed-torch field
        net.encode(io["images"], io["poses"], io["focal"], c=io["c"])
        oc = net(io["xyz"].requires_grad_(True), coarse=True, viewdirs=io["dirs"]).detach()
        of = net(io["xyz"], coarse=False, viewdirs=io["dirs"]).detach()
    assert (oc - io["out_coarse"]).abs().max() < 1e-5
    assert (of - io["out_fine"]).abs().max() < 1e-5
    # and back: our save_weights -> the reference's load_weights(strict=True) -> its forward gives the same values
    os.makedirs(os.path.join(d, "ours"), exist_ok=True)
    net.save_weights(Args(d, "ours"))
    net.save_weights(Args(d, "ours"))
    assert os.path.exists(os.path.join(d, "ours", "pixel_nerf_backup"))
    out = probe("load", d)
    assert float(out.strip().split("MAXDIFF")[-1]) < 1e-6
    # opt_init semantics (models.py:276-283): no resume + opt_init -> nothing is loaded, returns None
    assert net.load_weights(Args(d, "probe", resume=False), opt_init=True) is None

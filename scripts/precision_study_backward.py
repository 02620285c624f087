#!/usr/bin/env python
"""Which operand precision do the backward GEMMs need?  CPU emulation on the golden cases: every GEMM of
oracle/pnr_backward.py (dX = dY W, dW = dY^T X) is replaced by a reduced-precision product with exact (float64)
accumulation, and the resulting weight / latent gradients are compared with the fp32 ones.  Products of two fp16 or two
bf16 values are exact in fp32, so this isolates the operand rounding."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu  # noqa: E402

bw = gu.load_by_path("pnr_backward", os.path.join(ROOT, "oracle", "pnr_backward.py"))


def scaled(fn):
    """Per-tensor power-of-two scaling into the format's comfortable range (gradients are tiny)."""
    def q(x):
        m = float(x.abs().max())
        if m == 0.0:
            return x
        s = 2.0 ** (10 - torch.ceil(torch.log2(torch.tensor(m))).item())
        return fn(x * s) / s
    return q


f16 = scaled(lambda x: x.half().float())
bf16 = scaled(lambda x: x.bfloat16().float())


def hi_lo(q):
    def split(x):
        hi = q(x)
        return hi, q(x - hi)
    return split


MODES = {
    "fp32 (reference)": lambda a, b: a @ b,
    "fp16 x fp16, 1 product": lambda a, b: (f16(a).double() @ f16(b).double()).float(),
    "bf16 x bf16, 1 product": lambda a, b: (bf16(a).double() @ bf16(b).double()).float(),
    "fp16 hi/lo, 3 products": None,
    "bf16 hi/lo, 3 products": None,
}


def three(q):
    sp = hi_lo(q)

    def mm(a, b):
        ah, al = sp(a)
        bh, bl = sp(b)
        return (ah.double() @ bh.double() + al.double() @ bh.double() + ah.double() @ bl.double()).float()
    return mm


MODES["fp16 hi/lo, 3 products"] = three(f16)
MODES["bf16 hi/lo, 3 products"] = three(bf16)


def run(name):
    case = gu.load_case(name)
    cfg = case["cfg"]
    gt = torch.rand(cfg["SB"], cfg["B"], 3, generator=torch.Generator().manual_seed(9))
    res = {}
    for mode, mm in MODES.items():
        bw._mm = mm
        _, g_c, g_f, d_lat = bw.train_loss_backward(case["rays"], gt, case["noise"], gu.oracle_state(case),
                                                    case["latent"], case["wc"], case["wf"], cfg["NS"],
                                                    cfg["n_coarse"], cfg["n_fine"], cfg["n_fine_depth"],
                                                    white_bkgd=bool(cfg["white_bkgd"]))
        res[mode] = (g_c, g_f, d_lat)
    ref = res["fp32 (reference)"]
    print(f"case {name}: worst relative error (max|d| / max|ref|) over all weight tensors, and of the latent gradient")
    for mode, (g_c, g_f, d_lat) in res.items():
        worst = 0.0
        for gs, rs in ((g_c, ref[0]), (g_f, ref[1])):
            if gs is None:
                continue
            for k in gs:
                worst = max(worst, ((gs[k] - rs[k]).abs().max() / (rs[k].abs().max() + 1e-30)).item())
        lat = ((d_lat - ref[2]).abs().max() / ref[2].abs().max()).item()
        print(f"  {mode:28s} weights {worst:.2e}   latent {lat:.2e}")


for n in (sys.argv[1:] or ["sb2_d", "c2_small"]):
    run(n)

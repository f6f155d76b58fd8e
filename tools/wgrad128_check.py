"""A/B of the 128-byte-row weight-gradient kernel at the full-size layer shapes: repeatability, finiteness, and agreement
with the 64-byte-row kernel.  usage: wgrad128_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
B = 64
g = torch.Generator(device="cuda").manual_seed(0)
SHAPES = [(64, 0, 64, 128), (64, 0, 64, 64), (128, 0, 128, 64), (128, 64, 64, 64), (128, 0, 128, 32), (256, 128, 128, 32),
          (256, 0, 256, 16), (256, 256, 256, 16), (128, 0, 64, 64)]
for c0, c1, cout, r in SHAPES:
    x0 = torch.randn(B, r, r, c0, device="cuda", generator=g).bfloat16()
    x1 = torch.randn(B, r, r, c1, device="cuda", generator=g).bfloat16() if c1 else None
    dy = torch.randn(B, r, r, cout, device="cuda", generator=g).bfloat16()
    out = {}
    for opt in (0, 1):
        N.lib().pub_debug_option(b"wgrad_rows128", opt)
        runs = [N.conv2d_wgrad_nhwc(x0, dy, 3, x1=x1) for _ in range(4)]
        torch.cuda.synchronize()
        same = all(torch.equal(runs[0][0], r_[0]) and torch.equal(runs[0][1], r_[1]) for r_ in runs[1:])
        fin = bool(torch.isfinite(runs[0][0]).all()) and bool(torch.isfinite(runs[0][1]).all())
        out[opt] = runs[0]
        print(f"{c0}+{c1}->{cout}@{r}  rows128={opt} repeatable={same} finite={fin}", flush=True)
    dw = (out[0][0] - out[1][0]).abs().max() / out[0][0].abs().max()
    db = (out[0][1] - out[1][1]).abs().max() / out[0][1].abs().max()
    print(f"    rel diff dw {float(dw):.2e} db {float(db):.2e}", flush=True)

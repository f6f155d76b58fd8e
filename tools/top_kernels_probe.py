"""One launch each of the step's top conv kernels at their most frequent shapes (for `ncu --set full` captures).
usage: top_kernels_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
B = 64
g = torch.Generator(device="cuda").manual_seed(0)
for kind, cin, cout, r in [("wgrad", 32, 32, 128), ("wgrad", 256, 256, 16), ("stats", 32, 32, 128), ("stats", 256, 256, 16),
                           ("fwd", 64, 64, 64)]:
    x = torch.randn(B, r, r, cin, device="cuda", generator=g).bfloat16()
    if kind == "wgrad":
        dy = torch.randn(B, r, r, cout, device="cuda", generator=g).bfloat16()
        N.conv2d_wgrad_nhwc(x, dy, 3, want_bias=True)
    else:
        w = (torch.randn(9, cout, cin, device="cuda", generator=g) / (9 * cin) ** 0.5).bfloat16()
        y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.bfloat16)
        N.conv2d_nhwc(x, w, None, ksize=3, out=y, gn_stats=(kind == "stats"))
    torch.cuda.synchronize()
print("ok")

"""Runs a few representative conv launches (for `ncu --set full -k regex:...` captures and quick timing).
usage: conv_probe.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 64
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
SHAPES = [("fwd", 32, 32, 128), ("fwd", 64, 64, 64), ("fwd", 256, 256, 16), ("fwd", 128, 128, 32),
          ("wgrad", 32, 32, 128), ("wgrad", 64, 64, 64), ("wgrad", 256, 256, 16), ("wgrad", 128, 128, 32),
          ("wgrad", 64, 64, 128), ("wgrad", 96, 32, 128), ("wgrad", 128, 128, 64), ("wgrad", 512, 256, 16)]
for kind, cin, cout, r in SHAPES:
    x = torch.randn(B, r, r, cin, device="cuda", generator=g).bfloat16()
    fl = 2.0 * B * r * r * cin * cout * 9
    if kind == "fwd":
        w = torch.randn(9, cout, cin, device="cuda", generator=g).bfloat16()
        y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.bfloat16)
        fn = lambda: N.conv2d_nhwc(x, w, None, ksize=3, out=y)
    else:
        dy = torch.randn(B, r, r, cout, device="cuda", generator=g).bfloat16()
        fn = lambda: N.conv2d_wgrad_nhwc(x, dy, 3, want_bias=False)
    for box3 in ((0, 1) if kind == "wgrad" else (1,)):
        N.lib().pub_debug_option(b"wgrad_box3", box3)
        fn(); ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        tag = f" box3={box3}" if kind == "wgrad" else ""
        print(f"{kind:6s} {cin:4d}->{cout:4d} @{r:3d}^2  {t:8.1f} us  {fl / t / 1e6:7.1f} TFLOP/s{tag}", flush=True)

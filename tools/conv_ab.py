"""A/B timing of every distinct 3x3 tcgen05 conv launch of one training step: per-tap TMA kernel (conv_tc_kernel)
vs halo kernel (conv_halo_kernel), L2 flushed before each timed launch.  usage: conv_ab.py [B] [reps]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import _native as N
import bench
from helpers import canonical_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
TF32 = len(sys.argv) > 3 and sys.argv[3] == "tf32"     # the Gaussian encoders' dtype (f32 storage, kind::tf32)
# CONV_AB_ONLY="c0,c1,cout,res;..." restricts the sweep (ncu captures)
ONLY = [tuple(int(v) for v in it.split(",")) for it in os.environ.get("CONV_AB_ONLY", "").split(";") if it]
model = canonical_model(device="cuda")
inv = bench.conv_inventory(model, B, 128)
if TF32:   # encoder layers only: [3|6] -> 32 x3 @128, 64 x3 @64, 128 x3 @32, 256 x3 @16 (fwd + dgrad shapes)
    inv = {}
    for c, r in ((32, 128), (64, 64), (128, 32), (256, 16)):
        inv[("fwd", c, 0, c, r, 3)] = 8
        if c > 32:
            inv[("fwd", c // 2, 0, c, r, 3)] = 2
            inv[("fwd", c, 0, c // 2, r, 3)] = 2
cast = (lambda t: t.float()) if TF32 else (lambda t: t.bfloat16())
NDT = N.TF32 if TF32 else N.BF16
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
lib = N.lib()
# what is compared: CONV_AB_OPT="conv_halo:0,2" (default: per-tap vs halo everywhere) or e.g. "conv_chains:0,1"
OPT, VALS = os.environ.get("CONV_AB_OPT", "conv_halo:0,2").split(":")
VA, VB = (int(v) for v in VALS.split(","))
tot = {VA: 0.0, VB: 0.0}
rows = []
for (kind, c0, c1, cout, r, ks), cnt in sorted(inv.items()):
    if kind != "fwd" or ks != 3 or (c0 + c1) % 32 or cout % 32:
        continue
    if ONLY and (c0, c1, cout, r) not in ONLY:
        continue
    x0 = cast(torch.randn(B, r, r, c0, device="cuda", generator=g))
    x1 = cast(torch.randn(B, r, r, c1, device="cuda", generator=g)) if c1 else None
    w = cast(torch.randn(9, cout, c0 + c1, device="cuda", generator=g))
    y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.float32 if TF32 else torch.bfloat16)
    fl = 2.0 * B * r * r * (c0 + c1) * cout * 9
    res = {}
    for halo in (VA, VB):
        lib.pub_debug_option(OPT.encode(), halo)
        fn = lambda: N.conv2d_nhwc(x0, w, None, x1=x1, ksize=3, out=y, dtype=NDT)
        fn(); ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[halo] = statistics.median(ts)
        tot[halo] += res[halo] * cnt
    hbm = (B * r * r * (c0 + c1 + cout) * 2) / 6.5e6   # us at 6.5 TB/s (read x once, write y once)
    rows.append((res[VA] * cnt, f"{c0:4d}+{c1:<4d}->{cout:4d} @{r:3d}^2 x{cnt:<3d} {OPT}={VA} {res[VA]:7.1f} us {fl / res[VA] / 1e6:6.0f} TF | "
                 f"{OPT}={VB} {res[VB]:7.1f} us {fl / res[VB] / 1e6:6.0f} TF | hbm floor {hbm:6.1f} us  mma floor {fl / 2.25e9:6.1f} us"))
for _, line in sorted(rows, reverse=True):
    print(line)
lib.pub_debug_option(OPT.encode(), 1)
print(f"sum over one step: {OPT}={VA} {tot[VA] / 1e3:.3f} ms, {OPT}={VB} {tot[VB] / 1e3:.3f} ms")

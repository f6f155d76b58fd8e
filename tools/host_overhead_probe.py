"""Host-side enqueue cost of the engines (no device sync inside the timed regions)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import _native as N
from helpers import canonical_model
from climex_synth import make_fields
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
m = canonical_model(compute_dtype="bf16", device="cuda"); m.train()
f = make_fields(B, 128, 128, 16, seed=1)
x, y = f["inputs"].cuda(), f["targets"].cuda()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return dt * 1e3
with torch.no_grad():
    print("unet fwd enqueue ms", t(lambda: m.unet(x, _nhwc_out=True)))
    print("prior fwd enqueue ms", t(lambda: m.prior(x)))
l0 = N.lib().pub_launch_count()
with torch.no_grad(): m.unet(x, _nhwc_out=True)
print("unet fwd launches", N.lib().pub_launch_count() - l0)
def step():
    m.zero_grad(set_to_none=True)
    out = m.elbo(x, y, None, M=15)
    out[0].backward()
print("fwd+bwd enqueue ms", t(step, 3))
xb = torch.randn(B, 128, 128, 32, device="cuda").bfloat16()
w = torch.randn(9, 32, 32, device="cuda").bfloat16()
yb = torch.empty_like(xb)
print("conv2d_nhwc tc call us", t(lambda: N.conv2d_nhwc(xb, w, None, ksize=3, out=yb), 200) * 1e3)
print("conv2d_nhwc simt call us", t(lambda: N.conv2d_nhwc(xb, w, None, ksize=3, out=yb, backend=N.BACKEND_SIMT), 200) * 1e3)
wf = torch.randn(32, 32, 3, 3, device="cuda")
print("pack_conv_weight call us", t(lambda: N.pack_conv_weight(wf, N.BF16), 200) * 1e3)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

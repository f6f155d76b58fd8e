"""Looks for run-to-run differences of the weight-gradient kernel (dW and the fused bias sums) while another stream
keeps the SMs busy with the same kernel on other data -- the condition under which a lane-0-only stage release let the
TMA refill overtake late lanes (profiles/r02_wgrad_rows128.txt).  usage: wgrad_race.py [iterations]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
B = 64
ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator(device="cuda").manual_seed(0)
def mk(c, r): return torch.randn(B, r, r, c, device="cuda", generator=g).bfloat16()
side = torch.cuda.Stream()
for cin, cout, r in ((64, 64, 64), (32, 32, 128), (128, 128, 32)):
    x, dy, sx, sdy = mk(cin, r), mk(cout, r), mk(cin, r), mk(cout, r)
    ref = N.conv2d_wgrad_nhwc(x, dy, 3)
    torch.cuda.synchronize()
    bad_w = bad_b = 0
    for it in range(ITERS):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(6):
                N.conv2d_wgrad_nhwc(sx, sdy, 3)
        outs = [N.conv2d_wgrad_nhwc(x, dy, 3) for _ in range(4)]
        torch.cuda.synchronize()
        bad_w += sum(not torch.equal(o[0], ref[0]) for o in outs)
        bad_b += sum(not torch.equal(o[1], ref[1]) for o in outs)
    print(f"{cin}->{cout}@{r}^2 beside the same kernel on a second stream: differing dw {bad_w}/{4 * ITERS} db {bad_b}/{4 * ITERS}", flush=True)

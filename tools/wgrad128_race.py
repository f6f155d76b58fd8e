"""Looks for run-to-run differences of the 128-byte-row weight-gradient kernel while another stream keeps the SMs busy.
usage: wgrad128_race.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
B = 64
g = torch.Generator(device="cuda").manual_seed(0)
def mk(c, r): return torch.randn(B, r, r, c, device="cuda", generator=g).bfloat16()
x, dy = mk(64, 64), mk(64, 64)
side = torch.cuda.Stream()
LOADS = {
    "new-kernel wgrad 64->64@64": (mk(64, 64), mk(64, 64)),
}
ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ONLY = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (1, 0)
for rows128 in ONLY:
    N.lib().pub_debug_option(b"wgrad_rows128", rows128)
    N.lib().pub_debug_option(b"wgrad_rows128", rows128 & 1)
    ref = N.conv2d_wgrad_nhwc(x, dy, 3)
    N.lib().pub_debug_option(b"wgrad_rows128", rows128)
    torch.cuda.synchronize()
    for name, load in LOADS.items():
        bad_w = bad_b = 0
        shown = 0
        for it in range(ITERS):
            if load is not None:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(6):
                        N.conv2d_wgrad_nhwc(load[0], load[1], 3)
            outs = [N.conv2d_wgrad_nhwc(x, dy, 3) for _ in range(4)]
            torch.cuda.synchronize()
            bad_w += sum(not torch.equal(o[0], ref[0]) for o in outs)
            bad_b += sum(not torch.equal(o[1], ref[1]) for o in outs)
            for o in outs:
                if not torch.equal(o[1], ref[1]) and shown < 6:
                    shown += 1
                    d = (o[1] - ref[1])
                    nz = d.nonzero().flatten().tolist()
                    print("   db diff at channels", nz, "values", [round(float(d[i]), 3) for i in nz][:16], flush=True)
        print(f"rows128={rows128} side load: {name:32s} differing dw {bad_w}/{4 * ITERS} db {bad_b}/{4 * ITERS}", flush=True)

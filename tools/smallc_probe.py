"""First-layer (Cin = 3 / 6, 8-channel padded input) weight-gradient launches alone, for timing and
`ncu --set full -k regex:wgrad_smallc` captures.  usage: smallc_probe.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, R = 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for name, dt, ndt, cin in (("f32 cin6", torch.float32, N.F32, 6), ("f32 cin3", torch.float32, N.F32, 3),
                           ("bf16 cin3", torch.bfloat16, N.BF16, 3)):
    xp = torch.zeros(B, R, R, 8, device="cuda", dtype=dt)
    xp[..., :cin] = torch.randn(B, R, R, cin, device="cuda", generator=g).to(dt)
    x = xp[..., :cin]                      # channel slice of the padded staging buffer (ld = 8)
    dy = torch.randn(B, R, R, 32, device="cuda", generator=g).to(dt)
    fn = lambda: N.conv2d_wgrad_nhwc(x, dy, 3, want_bias=False, backend=N.BACKEND_SIMT, dtype=ndt)
    fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"wgrad_smallc {name}: {sorted(ts)[len(ts) // 2]:.1f} us (with its split-K finish)", flush=True)

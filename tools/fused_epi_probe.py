"""Timing + event trace of conv_halo_kernel with / without the fused GroupNorm epilogues (one launch, L2 flushed).
usage: fused_epi_probe.py c0 cout res [B] [mode: time|trace0|trace1|trace2|one1|one2]   (oneN: a single launch for ncu)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
c0, cout, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
mode = sys.argv[5] if len(sys.argv) > 5 else "time"
lib = N.lib()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, r, r, c0, device="cuda", generator=g).bfloat16()
w = (torch.randn(9, cout, c0, device="cuda", generator=g) / (9 * c0) ** 0.5).bfloat16()
y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.bfloat16)
gx = torch.randn(B, r, r, cout, device="cuda", generator=g).bfloat16()
coef = torch.randn(B, cout, 2, device="cuda", generator=g)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
bw = dict(x0=gx, coef=coef, p_drop=0.1, seed=5, subseq=2)
fns = {0: lambda: N.conv2d_nhwc(x, w, None, ksize=3, out=y),
       1: lambda: N.conv2d_nhwc(x, w, None, ksize=3, out=y, gn_stats=True),
       2: lambda: N.conv2d_nhwc(x, w, None, ksize=3, out=y, gn_bwd=bw),
       3: lambda: N.conv2d_nhwc(x, w, None, ksize=3, out=y, gn_bwd=dict(bw, p_drop=0.0))}
if mode.startswith("one"):
    fns[int(mode[3:])](); torch.cuda.synchronize(); sys.exit(0)
if mode == "time":
    for k, fn in fns.items():
        fn(); ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{c0}->{cout}@{r}^2 B={B} epi={k}: {sorted(ts)[2]:.1f} us   all {['%.1f' % t for t in ts]}")
    sys.exit(0)
k = int(mode[5:])
fns[k]()
tr = torch.zeros(3 * 1024, device="cuda", dtype=torch.int64)
lib.pub_debug_pointer(b"halo_trace", C.c_void_p(tr.data_ptr()))
flush.zero_(); torch.cuda.synchronize()
fns[k]()
torch.cuda.synchronize()
lib.pub_debug_pointer(b"halo_trace", C.c_void_p(0))
t = tr.cpu().view(3, 1024)
t0 = int(t[t > 0].min())
def rel(v): return [int(a) - t0 for a in v if a > 0]
P, M, E = rel(t[0]), rel(t[1]), rel(t[2])
cb = max(1, c0 // (64 if c0 % 64 == 0 else 32))
print(f"epi={k} shape {c0}->{cout}@{r}^2; K blocks per tile {cb}")
print("producer stamps:", P[:16])
per = 1 + 2 * cb
for i in range(0, min(len(M), per * 10), per):
    print("  mma tile", i // per, M[i:i + per])
print("epilogue group 0 (even tiles of the CTA): [acc full, stored]")
for i in range(0, min(len(E), 2 * 10), 2):
    print("  tile", i, E[i:i + 2], "dur", E[i + 1] - E[i] if i + 1 < len(E) else None)

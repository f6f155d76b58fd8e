"""Event timeline of CTA (0,0) of conv_halo_kernel (clock64 stamps): where does a K block spend its time?
usage: halo_trace.py c0 cout res [B]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200")):
    sys.path.insert(0, p)
import torch
import _native as N
c0, cout, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
lib = N.lib()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, r, r, c0, device="cuda", generator=g).bfloat16()
w = torch.randn(9, cout, c0, device="cuda", generator=g).bfloat16()
y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
N.conv2d_nhwc(x, w, None, ksize=3, out=y)
tr = torch.zeros(3 * 1024, device="cuda", dtype=torch.int64)
lib.pub_debug_pointer(b"halo_trace", C.c_void_p(tr.data_ptr()))
flush.zero_(); torch.cuda.synchronize()
N.conv2d_nhwc(x, w, None, ksize=3, out=y)
torch.cuda.synchronize()
lib.pub_debug_pointer(b"halo_trace", C.c_void_p(0))
t = tr.cpu().view(3, 1024)
t0 = int(t[t > 0].min())
cb = c0 // 32
def rel(v): return [int(a) - t0 for a in v if a > 0]
P, M, E = rel(t[0]), rel(t[1]), rel(t[2])
print(f"shape {c0}->{cout}@{r}^2  K blocks per tile {cb}; cycles relative to first stamp")
print("TMA producer, per K block: [stage free -> box issued]")
print("  ", P[:24])
print("mma: per tile [acc buffer free, then per K block: (A full, committed)]")
per = 1 + 2 * cb
for i in range(0, min(len(M), per * 8), per):
    print("  tile", i // per, M[i:i + per])
print("epilogue per tile: [acc full, stored]")
for i in range(0, min(len(E), 2 * 8), 2):
    print("  tile", i // 2, E[i:i + 2])
if len(P) >= 3:
    print("producer period (cycles/K block):", (P[-1] - P[1]) / max(1, len(P) - 2), "items", len(P))

"""cProfile of the HOST side of eager training steps (where do the ~21 ms of enqueue time per step go?).
usage: host_profile.py [steps]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import _native as N
from helpers import canonical_model
from climex_synth import make_fields
from optim import FusedAdamW
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, R, M = 64, 128, 15
model = canonical_model(latent_dim=32, loss_type="afcrps", compute_dtype="bf16", device="cuda")
model.train()
model.sync_scalars = False
N.manual_seed(1)
opt = FusedAdamW(model.parameters(), lr=1e-4)
f = make_fields(B, R, R, 16, seed=3)
x, y = f["inputs"].cuda(), f["targets"].cuda()
def step():
    opt.zero_grad(set_to_none=True)
    out = model.elbo(x, y, None, M=M)
    out[0].backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / steps:.2f} ms/step, wall {1e3 * (t2 - t0) / steps:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30)
print(s.getvalue()[:6000])

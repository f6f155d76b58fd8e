"""Does the captured training step replay?  usage: graph_probe.py [pdl 0|1] [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import _native as N
from helpers import canonical_model
from climex_synth import make_fields
from graph import GraphedTrainStep
from optim import FusedAdamW
pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N.lib().pub_debug_option(b"pdl", pdl)
m = canonical_model(compute_dtype="bf16", device="cuda"); m.train()
opt = FusedAdamW(m.parameters(), lr=1e-4)
f = make_fields(B, 128, 128, 16, seed=1)
x, y = f["inputs"].cuda(), f["targets"].cuda()
# eager timing first
m.sync_scalars = False
def eager():
    opt.zero_grad(set_to_none=True)
    out = m.elbo(x, y, None, M=15); out[0].backward(); opt.step(); return out
for _ in range(3): eager()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): eager()
torch.cuda.synchronize(); t_e = (time.perf_counter() - t0) / 10
g = GraphedTrainStep(m, opt, x, y, M=15, warmup=2)
print("captured: launches per step", g.launches_per_step, flush=True)
for _ in range(3): out = g(x, y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): out = g(x, y)
torch.cuda.synchronize(); t_g = (time.perf_counter() - t0) / 10
print(f"pdl={pdl} B={B}: eager {t_e*1e3:.2f} ms/step, graph {t_g*1e3:.2f} ms/step, loss {float(out[0]):.3f}, steps {int(g.counters[0])}")

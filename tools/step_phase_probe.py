"""Wall time of each phase of a training step over consecutive steps (sync after each phase)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import _native as N
from helpers import canonical_model
from climex_synth import make_fields
from optim import FusedAdamW
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = canonical_model(compute_dtype="bf16", device="cuda"); m.train()
opt = FusedAdamW(m.parameters(), lr=1e-4)
f = make_fields(B, 128, 128, 16, seed=1)
x, y = f["inputs"].cuda(), f["targets"].cuda()
sync = torch.cuda.synchronize
for it in range(8):
    T = []
    def mark(name, t0):
        sync(); T.append((name, (time.perf_counter() - t0) * 1e3))
    t0 = time.perf_counter(); opt.zero_grad(set_to_none=True); mark("zero", t0)
    t0 = time.perf_counter(); feat = m.unet(x, _nhwc_out=True); mark("unet", t0)
    t0 = time.perf_counter(); p = m.prior(x); q = m.posterior(x, y); mark("enc", t0)
    t0 = time.perf_counter(); kl = N.kl_normal(q.base_dist.loc, q.base_dist.scale, p.base_dist.loc, p.base_dist.scale); z = q.rsample((15,)); mark("latent", t0)
    t0 = time.perf_counter(); ens = N.fcomb_apply(m.fcomb, feat, z, nhwc=True); mark("fcomb", t0)
    t0 = time.perf_counter(); crps = N.ensemble_loss(ens, y, kind="afcrps"); total = crps + kl.mean(); mark("loss", t0)
    t0 = time.perf_counter(); total.backward(); mark("backward", t0)
    t0 = time.perf_counter(); opt.step(); mark("adamw", t0)
    print(it, " ".join(f"{n}={v:.1f}" for n, v in T), "sum=%.1f" % sum(v for _, v in T), flush=True)
print("mem allocated GB", torch.cuda.memory_allocated() / 2**30, "reserved GB", torch.cuda.memory_reserved() / 2**30)
print(torch.cuda.memory_stats().get("num_alloc_retries"), torch.cuda.memory_stats().get("num_device_alloc"), torch.cuda.memory_stats().get("num_device_free"))

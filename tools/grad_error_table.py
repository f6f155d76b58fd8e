"""Per-tensor element-wise gradient error of the CUDA path against the oracle (64x64, B=2, afCRPS M=3 and the
full-resolution B=16 case): the numbers the test tolerances in tests/test_gpu_model.py / test_gpu_fullsize.py are set from."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import canonical_model, rel_err
from oracle import probunet_oracle as O
g = np.load(os.path.join(ROOT, "tests", "golden", "probunet_golden.npz"))
x, y, eps = (torch.from_numpy(g[k]) for k in ("A_x", "A_y", "A_eps"))
cfg = O.ProbUNetCfg()
for name in ("fp32", "bf16"):
    m = canonical_model(compute_dtype=name, device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    O.elbo(full, cfg, x, y, eps, "afcrps")[0].backward()
    total, _, kl = m.elbo(x.cuda(), y.cuda(), None, M=3, eps=eps.cuda())
    total.backward()
    errs = sorted(((rel_err(p.grad, leaves[n].grad), n, float(leaves[n].grad.norm())) for n, p in m.named_parameters()
                   if leaves[n].grad is not None and float(leaves[n].grad.norm()) > 1e-7), reverse=True)
    print(f"== {name}: {len(errs)} tensors; median {errs[len(errs)//2][0]:.2e}; worst:")
    for e, n, nr in errs[:8]:
        print(f"   {e:.3e}  |g|={nr:.2e}  {n}")

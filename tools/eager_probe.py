"""Same-box PyTorch-eager baseline: the oracle's training step (functional torch restatement of the reference,
cuDNN / cuBLAS / ATen kernels) on cuda:0, fp32 (TF32 off), TF32 on, and bf16 autocast.  Informational."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from helpers import canonical_model
from oracle import probunet_oracle as O
from climex_synth import make_fields

def run(B, M, mode, steps=5, warmup=2, loss="afcrps", res=128, latent=32):
    m = canonical_model(latent_dim=latent)
    sd = {k: v.detach().clone().cuda() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4, fused=True)
    cfg = O.ProbUNetCfg(latent_dim=latent)
    f = make_fields(B, res, res, 16, seed=1237)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    enc, dec = O.unet_topology(cfg.unet())
    keys = [(b.key, b.cout, (b.up, b.down)) for b in enc + dec if not b.is_conv]
    g = torch.Generator(device="cuda").manual_seed(1)
    def step():
        eps = torch.randn(M, B, latent, device="cuda", generator=g)
        masks, h = {}, res
        for k, c, (up, down) in keys:
            h = h * 2 if up else (h // 2 if down else h)
            masks[k] = torch.rand(B, c, h, h, device="cuda", generator=g) >= 0.1
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            out = O.elbo(full, cfg, x, y, eps, loss, drop_masks=masks)
        out[0].backward()
        opt.step()
        return out[0]
    for _ in range(warmup): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode, "batch": B, "members": M, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "loss": float(l)}

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    for mode in ("fp32", "tf32", "bf16"):
        try:
            print(json.dumps(run(B, M, mode)), flush=True)
        except Exception as ex:
            print(json.dumps({"mode": mode, "error": repr(ex)[:300]}), flush=True)
        torch.cuda.empty_cache()

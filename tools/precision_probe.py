"""Prints the rel-err of every output of the bf16 / fp32 paths against the golden vectors (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from helpers import canonical_model, rel_err
import _native as N
g = np.load(os.path.join(ROOT, "tests/golden/probunet_golden.npz"))
x, y, eps = (torch.from_numpy(g[k]).cuda() for k in ("A_x", "A_y", "A_eps"))
for dt in ("fp32", "bf16"):
    m = canonical_model(compute_dtype=dt, device="cuda")
    with torch.no_grad():
        f = m.unet(x); p = m.prior(x); q = m.posterior(x, y)
        print(dt, "unet", rel_err(f, g["A_unet"]))
        print(dt, "prior mu", rel_err(p.base_dist.loc, g["A_prior_mu"]), "sigma", rel_err(p.base_dist.scale, g["A_prior_sigma"]),
              "log sigma abs", float((p.base_dist.scale.log().cpu() - torch.from_numpy(g["A_prior_sigma"]).log()).abs().max()))
        print(dt, "post mu", rel_err(q.base_dist.loc, g["A_post_mu"]), "sigma", rel_err(q.base_dist.scale, g["A_post_sigma"]),
              "log sigma abs", float((q.base_dist.scale.log().cpu() - torch.from_numpy(g["A_post_sigma"]).log()).abs().max()))
        print(dt, "  |mu| max", float(np.abs(g["A_post_mu"]).max()), "sigma range", float(g["A_post_sigma"].min()), float(g["A_post_sigma"].max()))
        zq = torch.from_numpy(g["A_post_mu"] + g["A_post_sigma"] * g["A_eps"][2]).cuda()
        print(dt, "fcomb(exact feat, exact z)", rel_err(m.fcomb(torch.from_numpy(g["A_unet"]).cuda(), zq), g["A_fcomb"]))
        print(dt, "fcomb(our feat, exact z)", rel_err(m.fcomb(f, zq), g["A_fcomb"]))
        z2 = q.base_dist.loc + q.base_dist.scale * eps[2]
        print(dt, "fcomb(exact feat, our z)", rel_err(m.fcomb(torch.from_numpy(g["A_unet"]).cuda(), z2), g["A_fcomb"]))
        print(dt, "fwd_train", rel_err(m(x, y, training=True, eps=eps[0]), g["A_fwd_train"]))
    m.loss_type = "afcrps"
    total, recon, kl = m.elbo(x, y, None, M=3, eps=eps)
    print(dt, "afcrps total", float(total), float(g["A_afcrps_total"]), "crps", recon[0], float(g["A_afcrps_crps"]), "kl", rel_err(kl, g["A_kl"]))

"""afCRPS / total-ELBO deviation of the bf16 path from the golden reference values under the A/B knobs
(conv_halo, fcomb_fwd_mma): which kernel choice moves the CRPS?  usage: crps_parity_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for q in (ROOT, ROOT + "/prob-unet-climate-downscaling_b200", ROOT + "/tests"):
    sys.path.insert(0, q)
import numpy as np, torch
import _native as N
from helpers import canonical_model, rel_err
g = np.load(ROOT + "/tests/golden/probunet_golden.npz")
x, y, eps = (torch.from_numpy(g[k]).cuda() for k in ("A_x", "A_y", "A_eps"))
for dt in ("bf16", "fp32"):
    m = canonical_model(compute_dtype=dt, device="cuda")
    for halo in (0, 1):
        for fm in (0, 1):
            N.lib().pub_debug_option(b"conv_halo", halo); N.lib().pub_debug_option(b"fcomb_fwd_mma", fm)
            with torch.no_grad():
                total, recon, kl = m.elbo(x, y, None, M=3, eps=eps)
                f = m.unet(x)
            e = abs(recon[0] - float(g["A_afcrps_crps"])) / float(g["A_afcrps_crps"])
            print(f"{dt} halo={halo} fcomb_mma={fm}: crps {recon[0]:.6f} (ref {float(g['A_afcrps_crps']):.6f}) rel dev {e:.2e}  unet feat rel_err {rel_err(f, g['A_unet']):.2e}")
N.lib().pub_debug_option(b"conv_halo", 1); N.lib().pub_debug_option(b"fcomb_fwd_mma", 1)

import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for q in (ROOT, ROOT+"/prob-unet-climate-downscaling_b200", ROOT+"/tests"): sys.path.insert(0, q)
import numpy as np, torch
from helpers import canonical_model
from oracle import probunet_oracle as O
import prob_unet_utils as U
g = np.load(ROOT+"/tests/golden/probunet_golden.npz")
x, y, eps = (torch.from_numpy(g[k]).cuda() for k in ("B_x", "B_y", "B_eps"))
m = canonical_model(compute_dtype="fp32", loss_type="mse+ssim", device="cuda")
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
with torch.no_grad():
    pred = m(x, y, training=True, eps=eps[0])
    ours = U.wmse_ms_ssim_loss(pred, y, return_components=True)
    ref = O.wmse_ms_ssim_loss(pred.cpu(), y.cpu())
    pref = O.forward(sd, O.ProbUNetCfg(), x.cpu(), y.cpu(), eps[0].cpu())
    ref2 = O.wmse_ms_ssim_loss(pref, y.cpu())
print("ours", [float(v) for v in ours]); print("oracle on our pred", [float(v) for v in ref]); print("oracle on oracle pred", [float(v) for v in ref2])
print("pred relerr", float((pred.cpu()-pref).norm()/pref.norm()), "golden recon", float(g["B_recon"]), float(g["B_msssim_loss"]), float(g["B_wmse"]))
print("pred stats", float(pred.std()), float(pred.abs().max()), "y range", float(y.max()-y.min()))
# per-level diagnostics from the oracle
X, Y = pred.cpu(), y.cpu()
import torch.nn.functional as F
win = O._gauss_1d(7, 1.5)
R = float((Y.max()-Y.min()).clamp(min=1e-5)); C1, C2 = (0.01*R)**2, (0.03*R)**2
for lvl in range(5):
    mu1, mu2 = O._gauss_filter(X, win), O._gauss_filter(Y, win)
    s11 = O._gauss_filter(X*X, win)-mu1*mu1; s22 = O._gauss_filter(Y*Y, win)-mu2*mu2; s12 = O._gauss_filter(X*Y, win)-mu1*mu2
    cs = ((2*s12+C2)/(s11+s22+C2)).flatten(2).mean(-1)
    print("lvl", lvl, "cs mean per plane", cs.flatten().tolist())
    X, Y = F.avg_pool2d(X, 2), F.avg_pool2d(Y, 2)

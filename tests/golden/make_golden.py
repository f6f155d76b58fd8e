"""Generate golden vectors from the REAL reference (run in the authoring container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/{prob_unet,networks,prob_unet_utils}.py unmodified (with
matplotlib / pytorch_msssim stubbed in sys.modules), builds the canonical
model with manual_seed(42), de-zeroes the zero-initialised tensors with manual_seed(43),
injects eps (seed 44) and dropout masks, and stores inputs + outputs + losses + KL +
gradient norms as small .npz fixtures beside this script.  /root/reference does not exist
on the GPU box: tests only read the .npz files.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"))
REF = os.environ.get("PROBUNET_REFERENCE", "/root/reference/src")


def import_reference():
    from oracle import probunet_oracle as O
    for name in ("matplotlib", "matplotlib.pyplot", "pytorch_msssim"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    def ms_ssim_stub(X, Y, data_range=255, size_average=True, win_size=11, **kw):
        assert size_average
        return O.ms_ssim(X, Y, data_range, win_size=win_size)
    sys.modules["pytorch_msssim"].ms_ssim = ms_ssim_stub
    sys.path.insert(0, REF)
    import prob_unet  # noqa
    import prob_unet_utils  # noqa
    sys.path.pop(0)
    return prob_unet, prob_unet_utils


sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import dezero  # noqa: E402


class EpsInjector:
    """Replaces the N(0,1) draw inside Normal.rsample by a queue of given tensors."""
    def __init__(self, eps_list):
        self.q = list(eps_list)

    def __enter__(self):
        import torch.distributions.normal as N
        self.N, self.orig = N, N._standard_normal
        N._standard_normal = lambda shape, dtype, device: self.q.pop(0).to(dtype).reshape(shape)
        return self

    def __exit__(self, *a):
        self.N._standard_normal = self.orig


class DropoutInjector:
    """Replaces F.dropout by recorded Bernoulli keep-masks from a seeded generator."""
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.masks = []

    def __enter__(self):
        import torch.nn.functional as F
        self.F, self.orig = F, F.dropout

        def drop(x, p=0.5, training=True, inplace=False):
            if not training or p == 0:
                return x
            m = torch.rand(x.shape, generator=self.g) >= p
            self.masks.append(m)
            return x * m.to(x.dtype) / (1.0 - p)
        F.dropout = drop
        return self

    def __exit__(self, *a):
        self.F.dropout = self.orig


def grads_summary(model):
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    return names, np.array(norms)


def grads_projection(model):
    """Two signed projections per gradient tensor (helpers.grad_projection): unlike a norm they change when a
    gradient is transposed, mirrored or permuted, so they pin every one of the 391 gradients element-wise."""
    from helpers import grad_projection
    return np.array([[0.0, 0.0] if p.grad is None else grad_projection(p.grad) for _, p in model.named_parameters()])


def main():
    torch.set_num_threads(os.cpu_count())
    prob_unet, utils = import_reference()
    from climex_synth import make_fields

    out = {}
    L = 32
    torch.manual_seed(42)
    model = prob_unet.ProbabilisticUNet(3, 3, L, [32, 64, 128, 256], 32, [1, 2, 4, 8], 1.0, 1.0, 0.0)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    out["sd_keys"] = np.array(list(sd0.keys()))
    out["sd_numel"] = np.array([v.numel() for v in sd0.values()])
    out["sd_sum"] = np.array([float(v.double().sum()) for v in sd0.values()])
    out["sd_abssum"] = np.array([float(v.double().abs().sum()) for v in sd0.values()])
    dezero(model)
    sd1 = model.state_dict()
    out["sd1_sum"] = np.array([float(v.double().sum()) for v in sd1.values()])
    out["sd1_abssum"] = np.array([float(v.double().abs().sum()) for v in sd1.values()])
    model.eval()

    # ---- case A: 64x64, B=2 (BASELINE config 1 shape, smaller batch) ----
    B, H, W, M = 2, 64, 64, 3
    f = make_fields(B, H, W, lowres_scale=8, seed=1234 + 1)
    x, y = f["inputs"], f["targets"]
    eps = torch.randn(M, B, L, generator=torch.Generator().manual_seed(44))
    out["A_x"], out["A_y"], out["A_eps"] = x.numpy(), y.numpy(), eps.numpy()

    with torch.no_grad():
        out["A_unet"] = model.unet(x).numpy()
        dp, dq = model.prior(x), model.posterior(x, y)
        out["A_prior_mu"], out["A_prior_sigma"] = dp.base_dist.loc.numpy(), dp.base_dist.scale.numpy()
        out["A_post_mu"], out["A_post_sigma"] = dq.base_dist.loc.numpy(), dq.base_dist.scale.numpy()
        with EpsInjector([eps[0]]):
            out["A_fwd_train"] = model(x, y, training=True).numpy()
        with EpsInjector([eps[1]]):
            out["A_fwd_prior"] = model(x, None, training=False).numpy()
        z = dq.base_dist.loc + dq.base_dist.scale * eps[2]
        out["A_fcomb"] = model.fcomb(model.unet(x), z).numpy()
        out["A_kl"] = torch.distributions.kl.kl_divergence(dq, dp).numpy()

    # afCRPS ELBO (src/prob_unet.py:273-317, commented variant) executed with the
    # reference's own sub-modules and loss function
    def elbo_afcrps(model, x, y, eps_list, alpha=0.95):
        feat = model.unet(x)
        pr, po = model.prior(x), model.posterior(x, y)
        with EpsInjector(eps_list):
            ens = torch.stack([model.fcomb(feat, po.rsample()) for _ in eps_list], dim=1)
        crps = utils.afcrps_loss(ens, y, alpha=alpha)
        kl = torch.distributions.kl.kl_divergence(po, pr)
        return model.beta_0 * crps + model.beta_1 * kl.mean(), crps, kl, ens

    def elbo_l1(model, x, y, e):
        feat = model.unet(x)
        pr, po = model.prior(x), model.posterior(x, y)
        with EpsInjector([e]):
            o = model.fcomb(feat, po.rsample())
        l1 = torch.nn.L1Loss()(o, y)
        kl = torch.distributions.kl.kl_divergence(po, pr)
        return model.beta_0 * l1 + model.beta_1 * kl.mean(), l1, kl

    model.zero_grad()
    total, crps, kl, ens = elbo_afcrps(model, x, y, [eps[m] for m in range(M)])
    total.backward()
    out["A_afcrps_total"], out["A_afcrps_crps"] = float(total), float(crps)
    out["A_afcrps_ens"] = ens.detach().numpy()
    out["A_crps_loss"] = float(utils.crps_loss(ens.detach(), y))
    names, norms = grads_summary(model)
    out["grad_names"], out["A_afcrps_gradnorm"] = np.array(names), norms
    out["A_afcrps_gradproj"] = grads_projection(model)
    gsd = dict(model.named_parameters())
    for k in ["fcomb.layers.0.weight", "fcomb.layers.4.bias", "posterior.conv_mu.weight",
              "prior.conv_log_sigma.bias", "unet.out_norm.weight", "unet.enc.64x64_block0.skip.weight",
              "unet.dec.128x128_block2.affine.bias", "unet.enc.128x128_conv.bias"]:
        out["A_afcrps_grad::" + k] = gsd[k].grad.numpy().copy()

    model.zero_grad()
    total, l1, kl = elbo_l1(model, x, y, eps[0])
    total.backward()
    out["A_l1_total"], out["A_l1_l1"] = float(total), float(l1)
    _, out["A_l1_gradnorm"] = grads_summary(model)
    out["A_l1_gradproj"] = grads_projection(model)

    # ---- case A-drop: train-mode dropout with injected masks (L1 ELBO) ----
    model.train()
    model.zero_grad()
    with DropoutInjector(45) as di:
        total, l1, kl = elbo_l1(model, x, y, eps[0])
    total.backward()
    model.eval()
    out["A_drop_l1_total"] = float(total)
    out["A_drop_nmask"] = len(di.masks)
    out["A_drop_maskbits"] = np.concatenate([np.packbits(m.numpy().reshape(-1)) for m in di.masks])
    out["A_drop_maskshapes"] = np.array([list(m.shape) for m in di.masks])
    _, out["A_drop_l1_gradnorm"] = grads_summary(model)
    out["A_drop_l1_gradproj"] = grads_projection(model)

    # ---- case B: active MS-SSIM ELBO (src/prob_unet.py:229-267) at 128x128, B=1 ----
    f = make_fields(1, 128, 128, lowres_scale=16, seed=1234 + 3)
    xb, yb = f["inputs"], f["targets"]
    eb = torch.randn(1, 1, L, generator=torch.Generator().manual_seed(46))
    model.zero_grad()
    with EpsInjector([eb[0]]):
        total, recon, klb, wmse, msl = model.elbo(xb, yb, None, M=1)
    total.backward()
    out["B_x"], out["B_y"], out["B_eps"] = xb.numpy(), yb.numpy(), eb.numpy()
    out["B_total"], out["B_recon"], out["B_kl"] = float(total), float(recon[0]), klb.detach().numpy()
    out["B_wmse"], out["B_msssim_loss"] = float(wmse), float(msl)
    _, out["B_gradnorm"] = grads_summary(model)
    out["B_gradproj"] = grads_projection(model)
    with torch.no_grad():
        fb = model.unet(xb)
        out["B_unet_sum"], out["B_unet_abssum"] = float(fb.double().sum()), float(fb.double().abs().sum())

    # ---- loss / metric known answers (in-tree reference functions) ----
    g = torch.Generator().manual_seed(47)
    e5 = torch.randn(2, 5, 3, 16, 16, generator=g)
    t5 = torch.randn(2, 3, 16, 16, generator=g)
    out["L_ens"], out["L_tgt"] = e5.numpy(), t5.numpy()
    out["L_afcrps"] = float(utils.afcrps_loss(e5, t5, alpha=0.95))
    out["L_crps"] = float(utils.crps_loss(e5, t5))
    sys.path.insert(0, REF)
    for name in ("climex_utils", "wandb", "xarray", "dask", "dask.distributed", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    try:
        import trainmodel
        ce = trainmodel.crps_empirical(e5.permute(1, 0, 2, 3, 4).contiguous(), t5)
        out["L_crps_empirical_mean"] = float(ce.mean())
    except Exception as ex:  # trainmodel drags in the data stack; crps_loss already pins CRPS
        print("trainmodel.crps_empirical not importable:", ex)
    np.savez_compressed(os.path.join(HERE, "probunet_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "probunet_golden.npz"),
          os.path.getsize(os.path.join(HERE, "probunet_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""CPU oracle for the Probabilistic U-Net hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (plain torch fp32 on CPU, functional over a
``state_dict``) of the reference algorithm in ``/root/reference/src``.  It is the
checker the GPU path is compared against; it is never imported by the product
package (``prob-unet-climate-downscaling_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference (``/root/reference/src/prob_unet.py``) in the authoring container and
stores its outputs / losses / KL / gradient norms under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those
vectors.  Third-party arithmetic that is *not* in the reference tree
(``pytorch_msssim.ms_ssim`` v1.0.0, ``pysteps...CRPS``) is restated from the
published algorithms and is "parity unpinned" for those two functions only
(see DESIGN.md); the CRPS restatement is cross-checked against the in-tree
``crps_loss`` / ``crps_empirical`` formulas which *are* pinned.

Citations ``file:line`` are relative to ``/root/reference/``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration / topology  (src/networks.py:226-297, src/prob_unet.py:146-189)
# --------------------------------------------------------------------------------------
@dataclass
class UNetCfg:
    in_channels: int = 3
    out_channels: int = 32
    model_channels: int = 32
    channel_mult: Sequence[int] = (1, 2, 4, 8)
    num_blocks: int = 2
    img_resolution: Tuple[int, int] = (128, 128)   # only used for the key names (src/networks.py:263-264)
    label_dim: int = 1
    dropout: float = 0.10


@dataclass
class BlockSpec:
    key: str            # e.g. "enc.64x64_block0"
    cin: int
    cout: int
    up: bool = False
    down: bool = False
    is_conv: bool = False   # the first encoder entry is a plain Conv2d (src/networks.py:269)


def unet_topology(cfg: UNetCfg) -> Tuple[List[BlockSpec], List[BlockSpec]]:
    """Restates the constructor loops of networks.UNet (src/networks.py:260-295)."""
    enc: List[BlockSpec] = []
    cout = cfg.in_channels
    for level, mult in enumerate(cfg.channel_mult):
        rx, ry = cfg.img_resolution[0] >> level, cfg.img_resolution[1] >> level
        if level == 0:
            cin, cout = cout, cfg.model_channels * mult
            enc.append(BlockSpec(f"enc.{rx}x{ry}_conv", cin, cout, is_conv=True))
        else:
            enc.append(BlockSpec(f"enc.{rx}x{ry}_down", cout, cout, down=True))
        for idx in range(cfg.num_blocks):
            cin, cout = cout, cfg.model_channels * mult
            enc.append(BlockSpec(f"enc.{rx}x{ry}_block{idx}", cin, cout))
    skips = [b.cout for b in enc]
    dec: List[BlockSpec] = []
    nlev = len(cfg.channel_mult)
    for level, mult in reversed(list(enumerate(cfg.channel_mult))):
        rx, ry = cfg.img_resolution[0] >> level, cfg.img_resolution[1] >> level
        if level == nlev - 1:
            dec.append(BlockSpec(f"dec.{rx}x{ry}_in0", cout, cout))
            dec.append(BlockSpec(f"dec.{rx}x{ry}_in1", cout, cout))
        else:
            dec.append(BlockSpec(f"dec.{rx}x{ry}_up", cout, cout, up=True))
        for idx in range(cfg.num_blocks + 1):
            cin = cout + skips.pop()
            cout = cfg.model_channels * mult
            dec.append(BlockSpec(f"dec.{rx}x{ry}_block{idx}", cin, cout))
    return enc, dec


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def gn_groups(c: int) -> int:
    """networks.GroupNorm: num_groups = min(32, C // 4)  (src/networks.py:100)."""
    return min(32, c // 4)


def group_norm(sd: SD, p: str, x: Tensor) -> Tensor:
    """src/networks.py:105-107 (eps 1e-5, affine)."""
    return F.group_norm(x, gn_groups(x.shape[1]), sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def resample(x: Tensor, up: bool, down: bool) -> Tensor:
    """The depthwise [1,1] box filter of networks.Conv2d (src/networks.py:64-66,83-87):
    ``up``  = conv_transpose2d(stride 2) with 4*f = ones(2,2)  == nearest 2x upsample,
    ``down`` = conv2d(stride 2) with f = 0.25*ones(2,2)          == 2x2 average pool."""
    if up:
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    if down:
        return F.avg_pool2d(x, 2)
    return x


def edm_conv(sd: SD, p: str, x: Tensor, up: bool = False, down: bool = False) -> Tensor:
    """networks.Conv2d.forward, non-fused branch (src/networks.py:82-91)."""
    x = resample(x, up, down)
    w = sd.get(p + ".weight")
    if w is not None:
        x = F.conv2d(x, w, padding=w.shape[-1] // 2)
        b = sd.get(p + ".bias")
        if b is not None:
            x = x + b.reshape(1, -1, 1, 1)
    return x


def unet_block(sd: SD, p: str, spec: BlockSpec, x: Tensor, film: Tensor,
               drop_mask: Optional[Tensor], p_drop: float) -> Tensor:
    """networks.UNetBlock.forward (src/networks.py:166-179), attention branch dead.

    ``film`` is ``affine(emb)``; with emb == 0 it is exactly ``affine.bias``
    (src/networks.py:310-314 feeds zeros through a bias-free map_label and silu).
    ``drop_mask`` (bool, same shape as the dropout input) injects the Bernoulli
    keep-mask of ``F.dropout`` (src/networks.py:177); None = no dropout (eval)."""
    orig = x
    h = edm_conv(sd, p + ".conv0", F.silu(group_norm(sd, p + ".norm0", x)), spec.up, spec.down)
    scale, shift = film.reshape(1, -1, 1, 1).chunk(2, dim=1)
    h = F.silu(torch.addcmul(shift, group_norm(sd, p + ".norm1", h), scale + 1))
    if drop_mask is not None:
        h = h * drop_mask.to(h.dtype) / (1.0 - p_drop)
    h = edm_conv(sd, p + ".conv1", h)
    has_skip = (spec.cin != spec.cout) or spec.up or spec.down      # src/networks.py:157
    skip = edm_conv(sd, p + ".skip", orig, spec.up, spec.down) if has_skip else orig
    return h + skip                                                # skip_scale == 1


def unet_forward(sd: SD, x: Tensor, cfg: UNetCfg, prefix: str = "unet.",
                 drop_masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """networks.UNet.forward (src/networks.py:299-333).  emb == silu(0) == 0."""
    enc, dec = unet_topology(cfg)
    skips: List[Tensor] = []
    for s in enc:
        p = prefix + s.key
        if s.is_conv:
            x = edm_conv(sd, p, x)
        else:
            m = None if drop_masks is None else drop_masks.get(s.key)
            x = unet_block(sd, p, s, x, sd[p + ".affine.bias"], m, cfg.dropout)
        skips.append(x)
    for s in dec:
        p = prefix + s.key
        if x.shape[1] != s.cin:
            x = torch.cat([x, skips.pop()], dim=1)
        m = None if drop_masks is None else drop_masks.get(s.key)
        x = unet_block(sd, p, s, x, sd[p + ".affine.bias"], m, cfg.dropout)
    x = F.silu(group_norm(sd, prefix + "out_norm", x))
    return edm_conv(sd, prefix + "out_conv", x)


def gaussian_encoder(sd: SD, p: str, x: Tensor, target: Optional[Tensor],
                     num_filters: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """AxisAlignedConvGaussian.forward (src/prob_unet.py:56-85) -> (mu, sigma)."""
    if target is not None:
        x = torch.cat([x, target], dim=1)
    idx = 0
    for i in range(len(num_filters)):
        if i != 0:
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        for _ in range(3):
            x = F.relu(F.conv2d(x, sd[f"{p}.encoder.{idx}.weight"], sd[f"{p}.encoder.{idx}.bias"], padding=1))
            idx += 2
    h = x.mean(dim=[2, 3], keepdim=True)
    mu = F.conv2d(h, sd[p + ".conv_mu.weight"], sd[p + ".conv_mu.bias"]).flatten(1)
    ls = F.conv2d(h, sd[p + ".conv_log_sigma.weight"], sd[p + ".conv_log_sigma.bias"]).flatten(1)
    return mu, torch.exp(ls) + 1e-7


def fcomb(sd: SD, feat: Tensor, z: Tensor, p: str = "fcomb") -> Tensor:
    """Fcomb.forward (src/prob_unet.py:120-138): tile z over H,W, concat, 3x 1x1 conv."""
    B, _, H, W = feat.shape
    zz = z[:, :, None, None].expand(B, z.shape[1], H, W)
    h = torch.cat([feat, zz], dim=1)
    h = F.relu(F.conv2d(h, sd[p + ".layers.0.weight"], sd[p + ".layers.0.bias"]))
    h = F.relu(F.conv2d(h, sd[p + ".layers.2.weight"], sd[p + ".layers.2.bias"]))
    return F.conv2d(h, sd[p + ".layers.4.weight"], sd[p + ".layers.4.bias"])


def kl_normal(mu_q: Tensor, sig_q: Tensor, mu_p: Tensor, sig_p: Tensor) -> Tensor:
    """kl_divergence(Independent(Normal q,1), Independent(Normal p,1)) -> [B]
    (torch.distributions.kl._kl_normal_normal; call site src/prob_unet.py:255)."""
    var_ratio = (sig_q / sig_p).pow(2)
    t1 = ((mu_q - mu_p) / sig_p).pow(2)
    return (0.5 * (var_ratio + t1 - 1 - var_ratio.log())).sum(-1)


# --------------------------------------------------------------------------------------
# losses  (src/prob_unet_utils.py)
# --------------------------------------------------------------------------------------
def afcrps_loss(ens: Tensor, target: Tensor, alpha: float = 0.95) -> Tensor:
    """src/prob_unet_utils.py:171-234, evaluated member-pair by member-pair so that no
    [B,M,M,C,H,W] temporary is needed.  ens [B,M,C,H,W], target [B,C,H,W]."""
    B, M, C, H, W = ens.shape
    eps = (1.0 - alpha) / M
    d = (ens - target.unsqueeze(1)).abs()                       # |x_j - y|
    total = ens.new_zeros(B)
    for j in range(M):
        for k in range(M):
            if j == k:
                continue
            comb = d[:, j] + d[:, k] - (1.0 - eps) * (ens[:, j] - ens[:, k]).abs()
            total = total + comb.sum(dim=(1, 2, 3))
    per_b = total / (2.0 * M * (M - 1)) / (C * H * W)
    return per_b.mean()


def crps_loss(ens: Tensor, target: Tensor) -> Tensor:
    """src/prob_unet_utils.py:237-268:  E|X-y| - 0.5 E|X-X'| (1/M^2 normalisation)."""
    B, M = ens.shape[:2]
    first = (ens - target.unsqueeze(1)).abs().mean(dim=1)
    second = torch.zeros_like(first)
    for j in range(M):
        second = second + (ens[:, j:j + 1] - ens).abs().sum(dim=1)
    return (first - 0.5 * second / (M * M)).mean()


def _gauss_1d(size: int, sigma: float) -> Tensor:
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gauss_filter(x: Tensor, win: Tensor) -> Tensor:
    C = x.shape[1]
    k = win.numel()
    x = F.conv2d(x, win.reshape(1, 1, k, 1).repeat(C, 1, 1, 1), groups=C)
    return F.conv2d(x, win.reshape(1, 1, 1, k).repeat(C, 1, 1, 1), groups=C)


def ms_ssim(X: Tensor, Y: Tensor, data_range: float, win_size: int = 7, sigma: float = 1.5) -> Tensor:
    """Restatement of pytorch_msssim.ms_ssim v1.0.0 (size_average=True, K=(0.01,0.03),
    5 default weights) -- THIRD-PARTY, not in the reference tree: parity unpinned.
    Call site: src/prob_unet_utils.py:297."""
    assert min(X.shape[-2:]) > (win_size - 1) * 2 ** 4
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=X.dtype)
    win = _gauss_1d(win_size, sigma).to(X.dtype)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mcs = []
    for lvl in range(5):
        mu1, mu2 = _gauss_filter(X, win), _gauss_filter(Y, win)
        s11 = _gauss_filter(X * X, win) - mu1 * mu1
        s22 = _gauss_filter(Y * Y, win) - mu2 * mu2
        s12 = _gauss_filter(X * Y, win) - mu1 * mu2
        cs_map = (2 * s12 + C2) / (s11 + s22 + C2)
        ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
        ssim_pc = ssim_map.flatten(2).mean(-1)
        cs = cs_map.flatten(2).mean(-1)
        if lvl < 4:
            mcs.append(F.relu(cs))
            pad = [s % 2 for s in X.shape[2:]]
            X = F.avg_pool2d(X, 2, padding=pad)
            Y = F.avg_pool2d(Y, 2, padding=pad)
    ssim_pc = F.relu(ssim_pc)
    vals = torch.stack(mcs + [ssim_pc], dim=0)                 # [5,B,C]
    return torch.prod(vals ** weights.view(-1, 1, 1), dim=0).mean()


def wmse_ms_ssim_loss(pred: Tensor, target: Tensor, alpha=0.007, beta=0.048, lam=0.0):
    """src/prob_unet_utils.py:270-305 -> (combined, wmse, 1-msssim)."""
    data_range = float((target.max() - target.min()).clamp(min=1e-5))
    w = torch.clamp(alpha * torch.exp(beta * target), max=1.0)
    wmse = (w * (pred - target).pow(2)).mean()
    ms = 1.0 - ms_ssim(pred, target, data_range)
    return lam * wmse + (1.0 - lam) * ms, wmse, ms


# --------------------------------------------------------------------------------------
# the model-level path  (src/prob_unet.py:194-381)
# --------------------------------------------------------------------------------------
@dataclass
class ProbUNetCfg:
    input_channels: int = 3
    num_classes: int = 3
    latent_dim: int = 32
    num_filters: Sequence[int] = (32, 64, 128, 256)
    model_channels: int = 32
    channel_mult: Sequence[int] = (1, 2, 4, 8)
    beta_0: float = 1.0
    beta_1: float = 1.0
    beta_2: float = 0.0

    def unet(self) -> UNetCfg:
        return UNetCfg(self.input_channels, self.num_filters[0], self.model_channels,
                       tuple(self.channel_mult))     # src/prob_unet.py:158-166


def forward(sd: SD, cfg: ProbUNetCfg, x: Tensor, target: Optional[Tensor], eps: Tensor,
            training: bool = True, drop_masks=None) -> Tensor:
    """ProbabilisticUNet.forward (src/prob_unet.py:194-224) with the N(0,1) draw of
    ``rsample`` injected as ``eps`` [B,L]."""
    feat = unet_forward(sd, x, cfg.unet(), drop_masks=drop_masks)
    if training and target is not None:
        mu, sig = gaussian_encoder(sd, "posterior", x, target, cfg.num_filters)
    else:
        mu, sig = gaussian_encoder(sd, "prior", x, None, cfg.num_filters)
    return fcomb(sd, feat, mu + sig * eps)


def elbo(sd: SD, cfg: ProbUNetCfg, x: Tensor, target: Tensor, eps: Tensor, loss_type: str,
         drop_masks=None, alpha: float = 0.95):
    """The three ELBO variants of the reference; ``eps`` is [M,B,L].

    * ``"afcrps"``  src/prob_unet.py:273-317  -> (total, crps, kl[B])
    * ``"l1"``      src/prob_unet.py:325-381  -> (total, l1_total, kl[B], kl2[B], l1_per_var[3])
    * ``"mse+ssim"`` src/prob_unet.py:229-267 -> (total, recon, kl[B], wmse, 1-msssim)
    """
    feat = unet_forward(sd, x, cfg.unet(), drop_masks=drop_masks)
    mu_p, sig_p = gaussian_encoder(sd, "prior", x, None, cfg.num_filters)
    mu_q, sig_q = gaussian_encoder(sd, "posterior", x, target, cfg.num_filters)
    kl = kl_normal(mu_q, sig_q, mu_p, sig_p)
    M = eps.shape[0]
    preds = [fcomb(sd, feat, mu_q + sig_q * eps[m]) for m in range(M)]
    if loss_type == "afcrps":
        assert M >= 2
        crps = afcrps_loss(torch.stack(preds, dim=1), target, alpha=alpha)
        return cfg.beta_0 * crps + cfg.beta_1 * kl.mean(), crps, kl
    if loss_type == "l1":
        out = preds[0]
        per_var = torch.stack([(out[:, i] - target[:, i]).abs().mean() for i in range(out.shape[1])])
        l1 = (out - target).abs().mean()
        kl2 = kl_normal(mu_q, sig_q, torch.zeros_like(mu_q), torch.ones_like(sig_q))
        total = cfg.beta_0 * l1 + cfg.beta_1 * kl.mean() + cfg.beta_2 * kl2.mean()
        return total, l1, kl, kl2, per_var
    if loss_type == "mse+ssim":
        recs = [wmse_ms_ssim_loss(p, target) for p in preds]
        recon = torch.stack([r[0] for r in recs]).mean()
        return cfg.beta_0 * recon + cfg.beta_1 * kl.mean(), recon, kl, recs[-1][1], recs[-1][2]
    raise ValueError(loss_type)


# --------------------------------------------------------------------------------------
# ensemble metrics  (src/metrics.py, src/trainmodel.py:66-110, src/climex_utils.py)
# --------------------------------------------------------------------------------------
def crps_empirical(pred: Tensor, truth: Tensor) -> Tensor:
    """src/trainmodel.py:66-110 (pyro's sort-based CRPS); pred [M,...], truth [...]."""
    M = pred.shape[0]
    if M == 1:
        return (pred[0] - truth).abs()
    s = pred.sort(dim=0).values
    diff = s[1:] - s[:-1]
    w = torch.arange(1, M, dtype=pred.dtype) * torch.arange(M - 1, 0, -1, dtype=pred.dtype)
    w = w.reshape(w.shape + (1,) * (diff.dim() - 1))
    return (s - truth).abs().mean(0) - (diff * w).sum(0) / M ** 2


def crps_hersbach_np(ens: np.ndarray, obs: np.ndarray) -> float:
    """pysteps.verification.probscores.CRPS restated (Hersbach 2000 alpha/beta
    decomposition) -- THIRD-PARTY, parity unpinned; call site src/metrics.py:39-41.
    ens [M,H,W], obs [H,W] -> scalar mean over finite pixels."""
    M = ens.shape[0]
    X = ens.reshape(M, -1).T.astype(np.float64)
    o = obs.reshape(-1).astype(np.float64)
    ok = np.isfinite(o) & np.all(np.isfinite(X), axis=1)
    X, o = np.sort(X[ok], axis=1), o[ok]
    n = X.shape[0]
    alpha = np.zeros((n, M + 1))
    beta = np.zeros((n, M + 1))
    for i in range(1, M):
        lo, hi = X[:, i - 1], X[:, i]
        a = np.where(o > hi, hi - lo, np.where(o > lo, o - lo, 0.0))
        b = np.where(o < lo, hi - lo, np.where(o < hi, hi - o, 0.0))
        alpha[:, i], beta[:, i] = a, b
    beta[:, 0] = np.where(o < X[:, 0], X[:, 0] - o, 0.0)
    alpha[:, M] = np.where(o > X[:, -1], o - X[:, -1], 0.0)
    p = np.arange(M + 1) / M
    return float(np.mean(np.sum(alpha * p ** 2 + beta * (1 - p) ** 2, axis=1)))


def crps_over_groundtruth(hr: Tensor, preds: Tensor) -> np.ndarray:
    """src/metrics.py:11-46 -> per-(t,var) CRPS array [T,3] (means are its column means)."""
    T, M, C, H, W = preds.shape
    out = np.zeros((T, C))
    for t in range(T):
        for v in range(C):
            out[t, v] = crps_hersbach_np(preds[t, :, v].numpy(), hr[t, v].numpy())
    return out


def compute_mae(hr: Tensor, preds: Tensor) -> np.ndarray:
    """src/metrics.py:48-71 -> per-(t,var) MAE of the ensemble mean, [T,3]."""
    pm = preds.mean(dim=1) if preds.dim() == 5 else preds
    return (hr - pm).abs().mean(dim=(2, 3)).numpy()


def softplus_ref(x: Tensor, threshold: float = 20.0, c: float = 1e-7) -> Tensor:
    """src/climex_utils.py:42-46."""
    return torch.where(x > threshold, x, torch.log(torch.exp(x) + 1.0) - c)


def invert_transfo_3vars(x: Tensor) -> Tensor:
    """results.ipynb cell 2 (the function cell 11 maps over every (t, member) before metrics.crps_over_groundtruth):
    pr = kgm2sTommday(softplus(x0)) (src/climex_utils.py:32-33,42-46), tasmin = KToC(x1) (:49-50),
    tasmax = KToC(softplus(x2) + x1) -- softplus with its default c = 1e-7 in both places.  [..., 3, H, W].
    Pinned by tests/golden/climex_golden.npz (tf_stored -> tf_real, produced by the notebook's own source)."""
    pr = softplus_ref(x[..., 0, :, :]) * 24 * 60 * 60          # kgm2sTommday multiplies in this order (fp32)
    tmin = x[..., 1, :, :] - 273.15
    tmax = (softplus_ref(x[..., 2, :, :]) + x[..., 1, :, :]) - 273.15
    return torch.stack([pr, tmin, tmax], dim=-3)


def residual_to_real(residual: Tensor, lrinterp: Tensor, std_hr: Tensor) -> Tensor:
    """residual_to_hr (src/climex_utils.py:277-285, epsilon 1e-10) followed by invert_transfo_3vars."""
    return invert_transfo_3vars(lrinterp + residual * (std_hr + 1e-10))


def psd_radial(image2d: Tensor):
    """results.ipynb cell 4 ``psd``: torch.fft.fftn power, radial wavenumber grid from fftfreq, mean per bin
    [k + 0.5, k + 1.5) (scipy.stats.binned_statistic 'mean': last edge closed) times the shell area."""
    H, W = image2d.shape
    power = (torch.abs(torch.fft.fftn(image2d)) ** 2).flatten().numpy().astype(np.float64)
    kfreq = torch.fft.fftfreq(H) * H
    kx, ky = torch.meshgrid(kfreq, kfreq, indexing="ij")
    kr = torch.sqrt(kx ** 2 + ky ** 2).flatten().numpy().astype(np.float64)
    kbins = np.arange(0.5, H // 2 + 1, 1.0)
    idx = np.searchsorted(kbins, kr, side="right") - 1
    idx[kr == kbins[-1]] = len(kbins) - 2
    ok = (idx >= 0) & (idx < len(kbins) - 1)
    sums = np.bincount(idx[ok], weights=power[ok], minlength=len(kbins) - 1)
    cnt = np.bincount(idx[ok], minlength=len(kbins) - 1)
    vals = sums / cnt * np.pi * (kbins[1:] ** 2 - kbins[:-1] ** 2)
    return 0.5 * (kbins[1:] + kbins[:-1]), vals


def compute_psd_tensor(data: Tensor, transfo: bool):
    """results.ipynb cell 4 ``compute_psd_tensor`` (tasmax: softplus with c = 0 in THIS cell)."""
    d = data.reshape(-1, *data.shape[-3:]) if data.dim() == 5 else data
    acc = [[], [], []]
    for smp in d:
        s0, s1, s2 = smp[0].clone(), smp[1].clone(), smp[2].clone()
        if transfo:
            s0 = softplus_ref(s0)
            s2 = softplus_ref(s2, c=0.0) + s1
        chans = (s0 * 24 * 60 * 60, s1 - 273.15, s2 - 273.15)
        for i in range(3):
            acc[i].append(psd_radial(chans[i])[1])
    return np.stack([np.mean(np.stack(a, axis=0), axis=0) for a in acc], axis=0)


def return_level_pixel_series(members_hr: Tensor, variable: str) -> Tensor:
    """test_return_levels.ipynb cell 2: the per-pixel daily value extracted from a residual_to_hr'd member
    (``members_hr`` [..., 3] = the three stored variables at the chosen pixel): pr = kgm2sTommday(softplus(x0)),
    tasmin = KToC(x1), tasmax = KToC(x1 + softplus(x2, c=0)) -- note c = 0 here, unlike results.ipynb cell 2."""
    if variable == "pr":
        return softplus_ref(members_hr[..., 0]) * 24 * 60 * 60
    if variable == "tasmin":
        return members_hr[..., 1] - 273.15
    if variable == "tasmax":
        return (members_hr[..., 1] + softplus_ref(members_hr[..., 2], c=0.0)) - 273.15
    raise ValueError("Unsupported variable")


# --------------------------------------------------------------------------------------
# dataset transform  (src/climex_utils.py:197-225, 255-264, 277-285)
# --------------------------------------------------------------------------------------
def climex_compute_stats(hr_all: Tensor, s: int):
    """compute_stats (src/climex_utils.py:255-264): statistics of the coarsened fields over time, and their
    block-constant expansion to the HR grid."""
    lr = F.avg_pool2d(hr_all, s)
    mean, std = lr.mean(dim=0), lr.std(dim=0)
    up = lambda t: t.repeat_interleave(s, dim=1).repeat_interleave(s, dim=2)
    return (mean, std), (up(mean), up(std))


def climex_getitem(hr: Tensor, stats, s: int, eps: float = 1e-10) -> Dict[str, Tensor]:
    """__getitem__ for type "lrinterp_to_residuals" (src/climex_utils.py:197-225), one field [C,H,W] or a batch."""
    batched = hr.dim() == 4
    h = hr if batched else hr.unsqueeze(0)
    lr = F.avg_pool2d(h, s)
    lrinterp = F.interpolate(lr, scale_factor=s)
    mean_hr, std_hr = stats[1]
    lrinterp_stand = (lrinterp - mean_hr) / (std_hr + eps)
    hr_stand = (h - mean_hr) / (std_hr + eps)
    out = {"inputs": lrinterp_stand, "targets": hr_stand - lrinterp_stand, "hr": h, "lr": lr, "lrinterp": lrinterp}
    return out if batched else {k: v.squeeze(0) for k, v in out.items()}

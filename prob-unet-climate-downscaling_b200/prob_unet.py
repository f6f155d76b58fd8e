"""Probabilistic U-Net -- host-side mirror of the reference's ``prob_unet.py`` module API.

Drop-in for /root/reference/src/prob_unet.py: same classes, constructor signature,
sub-module / parameter names (391-entry ``state_dict``), attributes and ``forward`` /
``elbo`` call conventions, so ``train_prob_unet_model.py`` and the latent-exploration
scripts run unchanged.  All arithmetic is done by hand-written sm_100a kernels behind
the C-ABI in ``include/probunet_b200.h``; there is no CPU fallback -- a missing
``libprobunet_b200.so`` or a CPU tensor raises.

Additive API (not in the reference): ``loss_type`` attribute selecting the ELBO variant
(SURVEY.md 3.4 B1), ``sample(x, n)`` (U-Net + prior once, ``fcomb`` n times) and
``compute_dtype``.
"""
import torch
import torch.nn as nn
from torch.distributions import Independent, Normal

import _native
from lazy_scalar import LazyScalar, lazy_list
from networks import UNet
from prob_unet_utils import init_weights, afcrps_loss, crps_loss, wmse_ms_ssim_loss  # noqa: F401

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


class LatentGaussian(Independent):
    """``Independent(Normal(loc, scale), 1)`` whose reparameterised draw is our Philox kernel.

    Keeps the torch.distributions surface the reference scripts use (``.base_dist.loc``,
    ``.base_dist.scale``, ``.rsample()``, ``kl_divergence``; src/latent_exploration_posterior.py:260-261)."""

    def __init__(self, loc, scale):
        super().__init__(Normal(loc=loc, scale=scale, validate_args=False), 1, validate_args=False)

    def rsample(self, sample_shape=torch.Size(), eps=None):
        if len(sample_shape) > 1:
            raise NotImplementedError("rsample supports () or (n,) sample shapes")
        n = sample_shape[0] if len(sample_shape) else None
        z = _native.rsample(self.base_dist.loc, self.base_dist.scale, n or 1, eps)
        return z if n is not None else z[0]

    def sample(self, sample_shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(sample_shape)


class AxisAlignedConvGaussian(nn.Module):
    """src/prob_unet.py:12-85: VGG-style conv encoder -> global mean -> (mu, log sigma) heads."""

    def __init__(self, input_channels, num_filters, latent_dim, posterior=False, compute_dtype=None):
        super().__init__()
        self.input_channels = input_channels
        self.num_filters = num_filters
        self.latent_dim = latent_dim
        self.posterior = posterior
        if self.posterior:
            self.input_channels += input_channels
        self.contracting_path = nn.ModuleList()
        layers, width = [], self.input_channels
        for i, nf in enumerate(self.num_filters):
            if i != 0:
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            for _ in range(3):
                layers.append(nn.Conv2d(width, nf, kernel_size=3, padding=1))
                layers.append(nn.ReLU(inplace=True))
                width = nf
        self.encoder = nn.Sequential(*layers)      # parameter container: indices 0,2,4,7,... as the reference
        self.conv_mu = nn.Conv2d(num_filters[-1], latent_dim, kernel_size=1)
        self.conv_log_sigma = nn.Conv2d(num_filters[-1], latent_dim, kernel_size=1)
        self.apply(init_weights)
        self.compute_dtype = compute_dtype
        self._engine = None

    def engine(self):
        dt = _native.resolve_encoder_dtype(self.compute_dtype)
        if self._engine is None or self._engine.dtype != dt:
            self._engine = _native.EncoderEngine(self, dt)
        return self._engine

    def forward(self, x, target=None):
        if not (self.posterior and target is not None):
            target = None
        mu, sigma = self.engine().forward(x, target)
        return LatentGaussian(mu, sigma)


class Fcomb(nn.Module):
    """src/prob_unet.py:87-138: broadcast z over H,W, concat with the features, 3x 1x1 conv."""

    def __init__(self, unet_output_channels, latent_dim, num_classes):
        super().__init__()
        self.latent_dim = latent_dim
        self.num_classes = num_classes
        self.channel_axis = 1
        self.spatial_axes = [2, 3]
        c = unet_output_channels
        self.layers = nn.Sequential(
            nn.Conv2d(c + latent_dim, c, kernel_size=1), nn.ReLU(inplace=True),
            nn.Conv2d(c, c, kernel_size=1), nn.ReLU(inplace=True),
            nn.Conv2d(c, num_classes, kernel_size=1))
        self.apply(init_weights)

    def tile(self, a, dim, n_tile):
        """TensorFlow-style tile == repeat_interleave along ``dim`` (src/prob_unet.py:109-118).
        Kept for the latent scripts (src/latent_exploration.py:524-525); not on the hot path."""
        return torch.repeat_interleave(a, n_tile, dim=dim)

    def forward(self, feature_map, z):
        """feature_map [B,F,H,W] (may be a non-contiguous expand() view), z [B,L] -> [B,C,H,W]."""
        return _native.fcomb_apply(self, feature_map, z.unsqueeze(0))[:, 0]

    def forward_members(self, feature_map, z):
        """z [M,B,L] -> [B,M,C,H,W]: one pass over the features for all M members."""
        return _native.fcomb_apply(self, feature_map, z)


class ProbabilisticUNet(nn.Module):
    """src/prob_unet.py:140-381."""

    def __init__(self, input_channels, num_classes, latent_dim, num_filters, model_channels, channel_mult,
                 beta_0, beta_1, beta_2, loss_type="afcrps", compute_dtype=None):
        super().__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.latent_dim = latent_dim
        self.model_channels = model_channels
        self.channel_mult = channel_mult
        self.beta_0, self.beta_1, self.beta_2 = beta_0, beta_1, beta_2
        self.loss_type = loss_type
        self.unet = UNet(img_resolution=(128, 128), in_channels=input_channels, out_channels=num_filters[0],
                         label_dim=1, model_channels=model_channels, channel_mult=channel_mult,
                         use_diffuse=False, compute_dtype=compute_dtype).to(device)
        self.prior = AxisAlignedConvGaussian(input_channels, num_filters, latent_dim, posterior=False,
                                             compute_dtype=compute_dtype).to(device)
        self.posterior = AxisAlignedConvGaussian(input_channels, num_filters, latent_dim, posterior=True,
                                                 compute_dtype=compute_dtype).to(device)
        self.fcomb = Fcomb(num_filters[0], latent_dim, num_classes).to(device)
        self.prior_latent_space = None
        self.posterior_latent_space = None
        # The reference returns the reconstruction terms as Python floats (.item() / .tolist(): a host sync between
        # forward and backward, measured at 1.1 ms of a 23.3 ms training step).  True: real floats, as the reference.
        # "lazy": float-like LazyScalars -- the copy to the host is enqueued where the reference blocks and awaited when
        # the value is first used (the reference's training loop appends them to a list and averages it per epoch).
        # False: 0-dim device tensors (CUDA-graph capture).  The values are the same in all three.
        self.sync_scalars = True

    # The three sub-networks are independent until fcomb / KL.  The U-Net's 32^2 and 16^2 levels and the encoders' late
    # stages launch fewer tiles than the GPU has SMs, so the two Gaussian encoders run on side streams next to the
    # U-Net (autograd replays that in backward: each backward node runs on its forward's stream).  Measured at B = 64:
    # 23.2 -> 22.5 ms per step (profiles/r02_encoder_streams_ab.txt).  PROBUNET_B200_ENCODER_STREAM: 0 off, 1 one side
    # stream, 3 (default) one per encoder, 2 / 4 the same with the U-Net enqueued first (slower).
    def _encoders_beside_unet(self, x, target):
        import os
        mode = int(os.environ.get("PROBUNET_B200_ENCODER_STREAM", "3"))
        if mode == 0 or not x.is_cuda or (torch.cuda.is_current_stream_capturing()
                                          and os.environ.get("PROBUNET_B200_GRAPH_STREAMS", "1") == "0"):
            feat = self.unet(x, _nhwc_out=True)
            prior = self.prior(x)
            post = self.posterior(x, target) if target is not None else None
            return feat, prior, post
        if getattr(self, "_side", None) is None:
            self._side = [torch.cuda.Stream(), torch.cuda.Stream()]
        cur = torch.cuda.current_stream()
        s_prior, s_post = self._side[0], self._side[1 if mode >= 3 else 0]
        unet_first = mode in (2, 4)          # experiment: autograd enqueues backward nodes in reverse creation order
        feat = self.unet(x, _nhwc_out=True) if unet_first else None
        s_prior.wait_stream(cur)
        if s_post is not s_prior:
            s_post.wait_stream(cur)
        with torch.cuda.stream(s_prior):
            prior = self.prior(x)
        post = None
        if target is not None:
            with torch.cuda.stream(s_post):
                post = self.posterior(x, target)
        if not unet_first:
            feat = self.unet(x, _nhwc_out=True)
        cur.wait_stream(s_prior)
        if s_post is not s_prior:
            cur.wait_stream(s_post)
        return feat, prior, post

    def set_compute_dtype(self, name):
        self.unet.compute_dtype = self.prior.compute_dtype = self.posterior.compute_dtype = name

    # src/prob_unet.py:194-224
    def forward(self, x, target=None, t=None, training=True, eps=None):
        feat = self.unet(x, _nhwc_out=True)
        if training and target is not None:
            self.posterior_latent_space = self.posterior(x, target)
            z = self.posterior_latent_space.rsample(eps=eps)
        else:
            self.prior_latent_space = self.prior(x)
            z = self.prior_latent_space.rsample(eps=eps)
        return _native.fcomb_apply(self.fcomb, feat, z.unsqueeze(0), nhwc=True)[:, 0]

    @torch.no_grad()
    def sample(self, x, n, eps=None):
        """n prior members per field: U-Net and prior run once, fcomb n times -> [B,n,C,H,W]."""
        feat, self.prior_latent_space, _ = self._encoders_beside_unet(x, None)
        z = self.prior_latent_space.rsample((n,), eps=eps)
        return _native.fcomb_apply(self.fcomb, feat, z, nhwc=True)

    @torch.no_grad()
    def sample_pixels(self, x, n, pixels, eps=None):
        """Additive API for the return-level analysis of test_return_levels.ipynb cell 2, which runs the full model
        ``num_samples`` times per day at batch 1 to read ONE pixel: here the U-Net and the prior run once per field and
        ``fcomb`` is evaluated only at the requested pixels (a gather of the [B,H,W,F] feature map in front of the
        per-pixel MLP -- fcomb has no spatial coupling, so this is exact).  ``pixels``: sequence of (y, x);
        returns the n prior members at those pixels, [B, n, C, P], in the model's (residual) domain."""
        ys = torch.as_tensor([int(p[0]) for p in pixels], device=x.device)
        xs = torch.as_tensor([int(p[1]) for p in pixels], device=x.device)
        if int(ys.max()) >= x.shape[2] or int(xs.max()) >= x.shape[3] or int(ys.min()) < 0 or int(xs.min()) < 0:
            raise IndexError("sample_pixels: pixel outside the field")
        feat = self.unet(x, _nhwc_out=True)                              # [B,H,W,F] engine layout
        self.prior_latent_space = self.prior(x)
        z = self.prior_latent_space.rsample((n,), eps=eps)
        fpix = feat[:, ys, xs, :].unsqueeze(1).contiguous()              # [B,1,P,F]: a 1 x P "image"
        return _native.fcomb_apply(self.fcomb, fpix, z, nhwc=True)[:, :, :, 0, :]

    @torch.no_grad()
    def sample_and_score(self, x, n, hr_real, lrinterp, std_hr, eps=None, field_batch=None, streams=2):
        """Additive API for the ensemble evaluation of results.ipynb cells 6, 11, 12: n prior members per field
        (``sample``), ``residual_to_hr`` + ``invert_transfo_3vars`` (src/climex_utils.py:277-285, results.ipynb cell 2)
        and ``metrics.crps_over_groundtruth`` / ``compute_mae`` (src/metrics.py:11-71) against ``hr_real`` [T,3,H,W]
        (real units) -> (crps [T,3], mae [T,3]) on the device; the members never leave the GPU.

        ``field_batch``: score the T fields in batches of that many (x / hr_real / lrinterp may then be pinned HOST
        tensors: each batch is copied on its own stream).  Batches are independent, so they alternate over ``streams``
        CUDA streams: the per-member kernels of one batch (fcomb, the sorting CRPS kernel -- few resident warps each)
        overlap the U-Net of the next (539 k -> 604 k members/s on one B200, profiles/r02_ensemble_streams_ab.txt)."""
        T = x.shape[0]
        dev = std_hr.device
        if field_batch is None or T <= field_batch:
            x, hr_real, lrinterp = (v.to(dev, non_blocking=True) for v in (x, hr_real, lrinterp))
            ens = self.sample(x, n, eps=eps)
            return _native.ensemble_metrics(ens, hr_real, lrinterp, std_hr)
        if eps is not None:
            raise ValueError("eps injection is per call: pass field_batch=None")
        crps = torch.empty(T, self.num_classes, device=dev, dtype=torch.float32)
        mae = torch.empty_like(crps)
        nst = max(1, int(streams))
        if getattr(self, "_batch_streams", None) is None or len(self._batch_streams) != nst:
            self._batch_streams = [torch.cuda.Stream(dev) for _ in range(nst)]
        cur = torch.cuda.current_stream(dev)
        for s_ in self._batch_streams:
            s_.wait_stream(cur)
        for k, i in enumerate(range(0, T, field_batch)):
            sl = slice(i, min(T, i + field_batch))
            with torch.cuda.stream(self._batch_streams[k % nst]):
                xb, hb, lb = (v[sl].to(dev, non_blocking=True) for v in (x, hr_real, lrinterp))
                c, a = _native.ensemble_metrics(self.sample(xb, n), hb, lb, std_hr)
                crps[sl], mae[sl] = c, a
        for s_ in self._batch_streams:
            cur.wait_stream(s_)
        return crps, mae

    def _host_scalar(self, v):
        """the reference's ``v.item()`` under the three ``sync_scalars`` policies"""
        if self.sync_scalars == "lazy":
            return LazyScalar(v)
        return v.detach().cpu().item() if self.sync_scalars else v.detach()

    def _host_list(self, v):
        """the reference's ``v.tolist()`` of a small 1-D tensor"""
        if self.sync_scalars == "lazy":
            return lazy_list(v)
        return v.detach().tolist() if self.sync_scalars else list(v.detach().unbind())

    def elbo(self, x, target, t=None, M=None, alpha=0.95, alpha_w=0.007, beta_w=0.048, lam_w=0.0, eps=None):
        """ELBO = beta_0*recon + beta_1*KL(q||p) [+ beta_2*KL(q||N(0,I))]; the reconstruction
        term and the return tuple follow ``self.loss_type`` exactly as the three variants in
        the reference (src/prob_unet.py:229-267 "mse+ssim", :273-317 "afcrps"/"crps", :325-381 "l1").
        ``eps`` [M,B,L] optionally injects the N(0,1) draws (parity tests)."""
        lt = self.loss_type
        if M is None:
            M = 5 if lt in ("afcrps", "crps") else 1
        if lt in ("afcrps", "crps") and M < 2:
            raise ValueError(f"M must be at least 2 to compute afCRPS but got M={M}")
        feat, self.prior_latent_space, self.posterior_latent_space = self._encoders_beside_unet(x, target)
        q, p = self.posterior_latent_space.base_dist, self.prior_latent_space.base_dist
        kl_div = _native.kl_normal(q.loc, q.scale, p.loc, p.scale)
        if lt == "l1":
            z = self.posterior_latent_space.rsample((1,), eps=eps)
            out = _native.fcomb_apply(self.fcomb, feat, z, nhwc=True)
            l1, per_var = _native.l1_loss(out[:, 0], target)
            kl2 = _native.kl_normal(q.loc, q.scale, torch.zeros_like(q.loc), torch.ones_like(q.scale))
            total = self.beta_0 * l1 + self.beta_1 * torch.mean(kl_div) + self.beta_2 * torch.mean(kl2)
            pv = per_var.detach()
            return total, self._host_list(pv), kl_div, kl2
        z = self.posterior_latent_space.rsample((M,), eps=eps)
        ens = _native.fcomb_apply(self.fcomb, feat, z, nhwc=True)      # [B,M,C,H,W]
        if lt in ("afcrps", "crps"):
            crps = _native.ensemble_loss(ens, target, kind=lt, alpha=alpha)
            total = self.beta_0 * crps + self.beta_1 * kl_div.mean()
            return total, [self._host_scalar(crps)], kl_div
        if lt == "mse+ssim":
            recs = [_native.wmse_ms_ssim(ens[:, m], target, alpha_w, beta_w, lam_w, None) for m in range(M)]
            recon = torch.stack([r[0] for r in recs]).mean()
            total = self.beta_0 * recon + self.beta_1 * kl_div.mean()
            host = self._host_scalar
            return total, [host(recon)], kl_div, host(recs[-1][1]), host(recs[-1][2])
        raise ValueError(f"unknown loss_type {lt!r}")

"""U-Net backbone -- host-side mirror of the reference's ``networks.py`` module API.

Same class names, constructor signatures, parameter/buffer names and initialisation
call order as /root/reference/src/networks.py (so ``torch.manual_seed(s)`` + construct
gives the reference's weights and reference ``.pth`` files load unchanged), but no
ATen compute: ``UNet.forward`` hands the whole network to the sm_100a engine in
``libprobunet_b200.so`` (see ``_native.py``).  There is no CPU fallback.
"""
import math

import torch

import _native


def weight_init(shape, mode, fan_in, fan_out):
    """Draws from the default CPU generator exactly like src/networks.py:21-26."""
    if mode == 'xavier_uniform':
        return math.sqrt(6 / (fan_in + fan_out)) * (torch.rand(*shape) * 2 - 1)
    if mode == 'xavier_normal':
        return math.sqrt(2 / (fan_in + fan_out)) * torch.randn(*shape)
    if mode == 'kaiming_uniform':
        return math.sqrt(3 / fan_in) * (torch.rand(*shape) * 2 - 1)
    if mode == 'kaiming_normal':
        return math.sqrt(1 / fan_in) * torch.randn(*shape)
    raise ValueError(f'Invalid init mode "{mode}"')


class Linear(torch.nn.Module):
    """src/networks.py:31-44.  On the hot path its input is identically zero (emb == 0),
    so the engine reads ``bias`` directly; ``weight`` only ever receives zero gradients."""

    def __init__(self, in_features, out_features, bias=True, init_mode='kaiming_normal', init_weight=1, init_bias=0):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        kw = dict(mode=init_mode, fan_in=in_features, fan_out=out_features)
        self.weight = torch.nn.Parameter(weight_init([out_features, in_features], **kw) * init_weight)
        self.bias = torch.nn.Parameter(weight_init([out_features], **kw) * init_bias) if bias else None

    def forward(self, x):
        raise RuntimeError("networks.Linear has no standalone kernel: it is folded into UNet.forward "
                           "(its input is identically zero on the Prob U-Net path)")


class Conv2d(torch.nn.Module):
    """src/networks.py:49-92: k x k same-padded conv with optional 2x box up/down resample
    (``kernel == 0`` -> resample only).  Parameter container; compute is in the engine."""

    def __init__(self, in_channels, out_channels, kernel, bias=True, up=False, down=False,
                 resample_filter=[1, 1], fused_resample=False, init_mode='kaiming_normal', init_weight=1, init_bias=0):
        assert not (up and down)
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.up, self.down, self.fused_resample = up, down, fused_resample
        self.kernel = kernel
        kw = dict(mode=init_mode, fan_in=in_channels * kernel * kernel, fan_out=out_channels * kernel * kernel)
        self.weight = torch.nn.Parameter(
            weight_init([out_channels, in_channels, kernel, kernel], **kw) * init_weight) if kernel else None
        self.bias = torch.nn.Parameter(weight_init([out_channels], **kw) * init_bias) if kernel and bias else None
        f = torch.as_tensor(resample_filter, dtype=torch.float32)
        f = f.ger(f).unsqueeze(0).unsqueeze(1) / f.sum().square()
        if (up or down) and list(resample_filter) != [1, 1]:
            raise NotImplementedError("only the [1,1] box resample filter used by the reference is implemented")
        self.register_buffer('resample_filter', f if up or down else None)

    def forward(self, x):
        if self.up or self.down or self.kernel not in (1, 3):
            raise RuntimeError("standalone networks.Conv2d with resampling is only available inside UNet.forward")
        return _native.conv2d(x, self.weight, self.bias)


class GroupNorm(torch.nn.Module):
    """src/networks.py:97-107 (groups = min(32, C // 4), eps 1e-5, affine)."""

    def __init__(self, num_channels, num_groups=32, min_channels_per_group=4, eps=1e-5):
        super().__init__()
        self.num_groups = min(num_groups, num_channels // min_channels_per_group)
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(num_channels))
        self.bias = torch.nn.Parameter(torch.zeros(num_channels))

    def forward(self, x):
        """Standalone F.group_norm (src/networks.py:105-107) on an NCHW tensor.  The engine only has the fused
        GroupNorm + SiLU kernel, so the plain normalisation is recovered as u = a*x + b from its per-(sample, channel)
        affine table; inside UNet.forward the fused path is used."""
        raise RuntimeError("standalone networks.GroupNorm is only available fused with SiLU: use "
                           "_native.groupnorm_silu_nhwc (the reference calls it only inside UNetBlock / UNet.forward)")


class UNetBlock(torch.nn.Module):
    """src/networks.py:134-187 without the (disabled) attention branch."""

    def __init__(self, in_channels, out_channels, emb_channels, up=False, down=False, attention=False,
                 num_heads=None, channels_per_head=64, dropout=0, skip_scale=1, eps=1e-5,
                 resample_filter=[1, 1], resample_proj=False, adaptive_scale=True,
                 init=dict(), init_zero=dict(init_weight=0), init_attn=None):
        super().__init__()
        if attention:
            raise NotImplementedError("attention is disabled everywhere in the reference (src/networks.py:275,285,295)")
        if not adaptive_scale or skip_scale != 1:
            raise NotImplementedError("only adaptive_scale=True, skip_scale=1 (the reference configuration)")
        self.in_channels, self.out_channels, self.emb_channels = in_channels, out_channels, emb_channels
        self.num_heads = 0
        self.dropout, self.skip_scale, self.adaptive_scale = dropout, skip_scale, adaptive_scale
        self.up, self.down = up, down
        self.norm0 = GroupNorm(num_channels=in_channels, eps=eps)
        self.conv0 = Conv2d(in_channels=in_channels, out_channels=out_channels, kernel=3, up=up, down=down,
                            resample_filter=resample_filter, **init)
        self.affine = Linear(in_features=emb_channels, out_features=out_channels * 2, **init)
        self.norm1 = GroupNorm(num_channels=out_channels, eps=eps)
        self.conv1 = Conv2d(in_channels=out_channels, out_channels=out_channels, kernel=3, **init_zero)
        self.skip = None
        if out_channels != in_channels or up or down:
            kernel = 1 if resample_proj or out_channels != in_channels else 0
            self.skip = Conv2d(in_channels=in_channels, out_channels=out_channels, kernel=kernel, up=up, down=down,
                               resample_filter=resample_filter, **init)

    def forward(self, x, emb):
        raise RuntimeError("standalone networks.UNetBlock is only available inside UNet.forward")


class UNet(torch.nn.Module):
    """src/networks.py:226-333.  ``forward(x)`` -> [B, out_channels, H, W] fp32 NCHW.

    Extra (ignored) label arguments are accepted so that both callers in the reference
    work: ``model.unet(x)`` (src/prob_unet.py:209) and ``model(inputs, class_labels=...)``
    (src/trainmodel.py:158).  ``compute_dtype`` ("bf16" | "fp32") selects the
    activation/tensor-core precision of the engine (fp32 master weights either way).
    """

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=1, augment_dim=0,
                 model_channels=16, channel_mult=[1, 4, 8, 16], channel_mult_emb=4, num_blocks=2,
                 attn_resolutions=[32, 16, 8], dropout=0.10, label_dropout=0, use_diffuse=False,
                 compute_dtype=None):
        super().__init__()
        if use_diffuse or augment_dim:
            raise NotImplementedError("use_diffuse / augment_dim are never enabled by the reference drivers")
        self.label_dropout = label_dropout
        self.dropout = dropout
        self.in_channels, self.out_channels = in_channels, out_channels
        emb_channels = model_channels * channel_mult_emb
        init = dict(init_mode='kaiming_uniform', init_weight=math.sqrt(1 / 3), init_bias=math.sqrt(1 / 3))
        init_zero = dict(init_mode='kaiming_uniform', init_weight=0, init_bias=0)
        block_kwargs = dict(emb_channels=emb_channels, channels_per_head=64, dropout=dropout, init=init, init_zero=init_zero)
        self.map_noise = None
        self.map_augment = None
        self.map_label = Linear(in_features=label_dim, out_features=emb_channels, bias=False,
                                init_mode='kaiming_normal', init_weight=math.sqrt(label_dim)) if label_dim else None
        assert len(img_resolution) == 2
        self.skips_postunet = None
        self.emb = None
        self.enc = torch.nn.ModuleDict()
        cout = in_channels
        for level, mult in enumerate(channel_mult):
            resx, resy = img_resolution[0] >> level, img_resolution[1] >> level
            if level == 0:
                cin, cout = cout, model_channels * mult
                self.enc[f'{resx}x{resy}_conv'] = Conv2d(in_channels=cin, out_channels=cout, kernel=3, **init)
            else:
                self.enc[f'{resx}x{resy}_down'] = UNetBlock(in_channels=cout, out_channels=cout, down=True, **block_kwargs)
            for idx in range(num_blocks):
                cin, cout = cout, model_channels * mult
                self.enc[f'{resx}x{resy}_block{idx}'] = UNetBlock(in_channels=cin, out_channels=cout, **block_kwargs)
        skips = [b.out_channels for b in self.enc.values()]
        self.dec = torch.nn.ModuleDict()
        for level, mult in reversed(list(enumerate(channel_mult))):
            resx, resy = img_resolution[0] >> level, img_resolution[1] >> level
            if level == len(channel_mult) - 1:
                self.dec[f'{resx}x{resy}_in0'] = UNetBlock(in_channels=cout, out_channels=cout, **block_kwargs)
                self.dec[f'{resx}x{resy}_in1'] = UNetBlock(in_channels=cout, out_channels=cout, **block_kwargs)
            else:
                self.dec[f'{resx}x{resy}_up'] = UNetBlock(in_channels=cout, out_channels=cout, up=True, **block_kwargs)
            for idx in range(num_blocks + 1):
                cin = cout + skips.pop()
                cout = model_channels * mult
                self.dec[f'{resx}x{resy}_block{idx}'] = UNetBlock(in_channels=cin, out_channels=cout, **block_kwargs)
        self.out_norm = GroupNorm(num_channels=cout)
        self.out_conv = Conv2d(in_channels=cout, out_channels=out_channels, kernel=3, **init_zero)
        self.compute_dtype = compute_dtype
        self._engine = None

    # -- engine plumbing --------------------------------------------------------------
    def engine(self):
        dt = _native.resolve_dtype(self.compute_dtype)
        if self._engine is None or self._engine.dtype != dt:
            self._engine = _native.UNetEngine(self, dt)
        return self._engine

    def forward(self, x, noise_labels=None, class_labels=None, augment_labels=None, _nhwc_out=False):
        return self.engine().forward(x, self.training, nhwc_out=_nhwc_out)

"""``torch.library`` registration of the C-ABI entry points: ``torch.ops.probunet_b200.*``.

BASELINE.json's north_star asks for "a thin C-ABI extension registered as torch custom ops".  The module mirror
(`prob_unet.py`, `networks.py`) drives whole sub-networks through `torch.autograd.Function`s over pointer tables
(`_native.py`: ~12 autograd nodes per training step); this file registers the *tensor-signature* entry points of
``include/probunet_b200.h`` with the dispatcher as well, each with a fake (meta) kernel and an autograd formula, so
that they compose with `torch.compile` / `torch.export` graphs and `torch.library.opcheck`:

    conv2d_nhwc(x, weight, bias?, relu)          pub_conv2d_forward  + pub_pack_conv_weight   (src/networks.py:89)
    conv2d_nhwc_backward(dy, x, weight, ...)     pub_conv2d_forward (mirrored weights) + pub_conv2d_wgrad
    fcomb(feat, z, w0, b0, w1, b1, w2, b2)       pub_fcomb_forward / pub_fcomb_backward       (src/prob_unet.py:120-138)
    ensemble_loss(ens, target, kind, alpha)      pub_ensemble_loss                            (src/prob_unet_utils.py:171-268)
    wmse_msssim(pred, target, alpha, beta, lam)  pub_wmse_msssim_loss                         (src/prob_unet_utils.py:270-305)
    kl_normal(mq, sq, mp, sp)                    pub_kl_normal_forward / _backward            (src/prob_unet.py:255)

CUDA only: there is no CPU kernel behind any of them (a CPU tensor raises NativeError).
"""
import ctypes as C

import torch

import _native as N

_NS = "probunet_b200"


def _dt(x):
    return N.BF16 if x.dtype == torch.bfloat16 else N.F32


# ------------------------------------------------------------------------------------------------ conv
@torch.library.custom_op(f"{_NS}::conv2d_nhwc", mutates_args=(), device_types="cuda")
def conv2d_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, relu: bool) -> torch.Tensor:
    """x [B,H,W,Cin] (bf16 or f32, NHWC), weight [Cout,Cin,k,k] f32 OIHW (k = 1 or 3, same padding) -> [B,H,W,Cout]."""
    wp = N.pack_conv_weight(weight, _dt(x))
    return N.conv2d_nhwc(x, wp, bias, relu=relu, ksize=weight.shape[-1])


@conv2d_nhwc.register_fake
def _(x, weight, bias, relu):
    return x.new_empty(x.shape[0], x.shape[1], x.shape[2], weight.shape[0])


@torch.library.custom_op(f"{_NS}::conv2d_nhwc_backward", mutates_args=(), device_types="cuda")
def conv2d_nhwc_backward(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, y: torch.Tensor | None,
                         need_dx: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dx, dweight, dbias).  ``y`` (the forward output) is passed when the forward applied ReLU."""
    k = weight.shape[-1]
    dy = dy.contiguous()
    if y is not None:
        dy = dy * (y > 0).to(dy.dtype)
    dw, db = N.conv2d_wgrad_nhwc(x.contiguous(), dy, k)
    if need_dx:
        dx = N.conv2d_nhwc(dy, N.pack_conv_weight(weight, _dt(x), transpose_flip=True), None, ksize=k)
    else:
        dx = x.new_zeros(())
    return dx, dw, db


@conv2d_nhwc_backward.register_fake
def _(dy, x, weight, y, need_dx):
    return (torch.empty_like(x) if need_dx else x.new_empty(()), weight.new_empty(weight.shape),
            weight.new_empty(weight.shape[0]))


def _conv_setup(ctx, inputs, output):
    x, weight, bias, relu = inputs
    ctx.relu, ctx.has_bias = relu, bias is not None
    ctx.save_for_backward(x, weight, output if relu else None)


def _conv_bwd(ctx, dy):
    x, weight, y = ctx.saved_tensors
    dx, dw, db = conv2d_nhwc_backward(dy, x, weight, y, ctx.needs_input_grad[0])
    return (dx if ctx.needs_input_grad[0] else None, dw, db if ctx.has_bias else None, None)


conv2d_nhwc.register_autograd(_conv_bwd, setup_context=_conv_setup)


# ------------------------------------------------------------------------------------------------ fcomb
class _FcombShim:
    """The C struct wants module-shaped access (layers[0/2/4], num_classes); wrap loose tensors."""

    class _L:
        def __init__(self, w, b):
            self.weight, self.bias = w, b

    def __init__(self, w0, b0, w1, b1, w2, b2):
        self.layers = [self._L(w0, b0), None, self._L(w1, b1), None, self._L(w2, b2)]
        self.num_classes = w2.shape[0]


@torch.library.custom_op(f"{_NS}::fcomb", mutates_args=(), device_types="cuda")
def fcomb(feat: torch.Tensor, z: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
          w2: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """feat [B,F,H,W] f32 (any strides), z [M,B,L] -> [B,M,C,H,W]: all M members in one pass over the features."""
    mod = _FcombShim(w0, b0, w1, b1, w2, b2)
    with torch.no_grad():
        return N._FcombFn.apply(mod, False, feat, z, w0, b0, w1, b1, w2, b2)


@fcomb.register_fake
def _(feat, z, w0, b0, w1, b1, w2, b2):
    return feat.new_empty(feat.shape[0], z.shape[0], w2.shape[0], feat.shape[2], feat.shape[3], dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::fcomb_backward", mutates_args=(), device_types="cuda")
def fcomb_backward(dout: torch.Tensor, feat: torch.Tensor, z: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor,
                   w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
                   b2: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                              torch.Tensor, torch.Tensor, torch.Tensor]:
    mod = _FcombShim(w0, b0, w1, b1, w2, b2)
    feat, z = feat.float(), z.contiguous().float()
    a = N._fcomb_args(mod, feat, z, False, None)
    nbytes = N.lib().pub_fcomb_backward_workspace(C.byref(a))
    ws = torch.empty(nbytes, device=z.device, dtype=torch.uint8)
    g = [torch.empty_like(p, dtype=torch.float32) for p in (w0, b0, w1, b1, w2, b2)]
    dz, dfeat = torch.empty_like(z), torch.empty(feat.shape, device=feat.device, dtype=torch.float32)
    N.check(N.lib().pub_fcomb_backward(C.byref(a), N.ptr(dout.contiguous()), N.ptr(dfeat), N.ptr(dz), *[N.ptr(t) for t in g],
                                       N.ptr(ws), C.c_size_t(nbytes), N.stream()), "pub_fcomb_backward")
    return (dfeat, dz, *g)


@fcomb_backward.register_fake
def _(dout, feat, z, w0, b0, w1, b1, w2, b2):
    return (feat.new_empty(feat.shape, dtype=torch.float32), torch.empty_like(z), torch.empty_like(w0), torch.empty_like(b0),
            torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2))


fcomb.register_autograd(lambda ctx, dout: fcomb_backward(dout, *ctx.saved_tensors),
                        setup_context=lambda ctx, inputs, output: ctx.save_for_backward(*inputs))


# ------------------------------------------------------------------------------------------------ losses
@torch.library.custom_op(f"{_NS}::ensemble_loss", mutates_args=(), device_types="cuda")
def ensemble_loss(ens: torch.Tensor, target: torch.Tensor, kind: int, alpha: float) -> tuple[torch.Tensor, torch.Tensor]:
    """ens [B,M,C,H,W], target [B,C,H,W]; kind 0 = afCRPS, 1 = CRPS -> (loss, d loss / d ens)."""
    B, M, Cc, H, W = ens.shape
    ens, target = ens.contiguous().float(), target.contiguous().float()
    loss = torch.empty((), device=ens.device, dtype=torch.float32)
    dens = torch.empty_like(ens)
    ws, n = N._loss_ws(B, Cc, H * W, ens.device)
    N.check(N.lib().pub_ensemble_loss(N.ptr(ens), N.ptr(target), B, M, Cc, H * W, kind, C.c_float(alpha), N.ptr(loss),
                                      N.ptr(dens), N.ptr(ws), C.c_size_t(n), N.stream()), "pub_ensemble_loss")
    return loss, dens


@ensemble_loss.register_fake
def _(ens, target, kind, alpha):
    return ens.new_empty((), dtype=torch.float32), torch.empty_like(ens, dtype=torch.float32)


def _ens_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.mark_non_differentiable(output[1])


ensemble_loss.register_autograd(lambda ctx, dl, _d: (ctx.saved_tensors[0] * dl, None, None, None), setup_context=_ens_setup)


@torch.library.custom_op(f"{_NS}::wmse_msssim", mutates_args=(), device_types="cuda")
def wmse_msssim(pred: torch.Tensor, target: torch.Tensor, alpha: float, beta: float,
                lam: float) -> tuple[torch.Tensor, torch.Tensor]:
    """-> ([loss, wmse, 1 - msssim], d loss / d pred)."""
    B, Cc, H, W = pred.shape
    pred, target = pred.contiguous().float(), target.contiguous().float()
    out3 = torch.empty(3, device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred)
    n = N.lib().pub_msssim_workspace(B, Cc, H, W)
    if n == 0:
        raise N.NativeError(N.lib().pub_last_error().decode())
    ws = torch.empty(n, device=pred.device, dtype=torch.uint8)
    N.check(N.lib().pub_wmse_msssim_loss(N.ptr(pred), N.ptr(target), B, Cc, H, W, C.c_float(alpha), C.c_float(beta),
                                         C.c_float(lam), N.ptr(out3), N.ptr(dpred), N.ptr(ws), C.c_size_t(n), N.stream()),
            "pub_wmse_msssim_loss")
    return out3, dpred


@wmse_msssim.register_fake
def _(pred, target, alpha, beta, lam):
    return pred.new_empty(3, dtype=torch.float32), torch.empty_like(pred, dtype=torch.float32)


def _ms_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])
    ctx.mark_non_differentiable(output[1])


wmse_msssim.register_autograd(lambda ctx, d3, _d: (ctx.saved_tensors[0] * d3[0], None, None, None, None), setup_context=_ms_setup)


# ------------------------------------------------------------------------------------------------ KL
@torch.library.custom_op(f"{_NS}::kl_normal", mutates_args=(), device_types="cuda")
def kl_normal(mq: torch.Tensor, sq: torch.Tensor, mp: torch.Tensor, sp: torch.Tensor) -> torch.Tensor:
    """KL( N(mq, sq) || N(mp, sp) ) summed over the latent dimension: [B,L] x4 -> [B]."""
    with torch.no_grad():
        return N.kl_normal(mq, sq, mp, sp)


@kl_normal.register_fake
def _(mq, sq, mp, sp):
    return mq.new_empty(mq.shape[0])


def _kl_bwd(ctx, dkl):
    mq, sq, mp, sp = ctx.saved_tensors
    with torch.enable_grad():
        leaves = [t.detach().requires_grad_(True) for t in (mq, sq, mp, sp)]
        out = N.kl_normal(*leaves)
    return torch.autograd.grad(out, leaves, dkl)


kl_normal.register_autograd(_kl_bwd, setup_context=lambda ctx, inputs, output: ctx.save_for_backward(*inputs))

OPS = ("conv2d_nhwc", "conv2d_nhwc_backward", "fcomb", "fcomb_backward", "ensemble_loss", "wmse_msssim", "kl_normal")

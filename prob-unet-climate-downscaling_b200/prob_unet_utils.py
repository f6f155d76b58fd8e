"""Initialisation + loss entry points -- host mirror of the reference's ``prob_unet_utils.py``
(only the hot-path part: src/prob_unet_utils.py:10-23 and :171-305).  The losses run as
sm_100a kernels (``csrc/loss.cu``) with hand-written backward; GEV / plotting helpers of
the reference are out of scope.
"""
import torch
import torch.nn as nn

import _native


def truncated_normal_(tensor, mean=0, std=1):
    """First of four N(0,1) candidates with |x| < 2 (src/prob_unet_utils.py:10-16)."""
    cand = tensor.new_empty(tuple(tensor.shape) + (4,)).normal_()
    ok = (cand < 2) & (cand > -2)
    first = ok.max(-1, keepdim=True)[1]
    tensor.data.copy_(cand.gather(-1, first).squeeze(-1))
    tensor.data.mul_(std).add_(mean)


def init_weights(m):
    """kaiming-normal weights, 1e-3 * truncated-normal bias (src/prob_unet_utils.py:18-23)."""
    if type(m) == nn.Conv2d or type(m) == nn.ConvTranspose2d:
        nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
        truncated_normal_(m.bias, mean=0, std=0.001)


def afcrps_loss(ensemble_pred, target, alpha=0.95):
    """almost-fair CRPS, [B,M,C,H,W] x [B,C,H,W] -> scalar (src/prob_unet_utils.py:171-234)."""
    return _native.ensemble_loss(ensemble_pred, target, kind="afcrps", alpha=alpha)


def crps_loss(ensemble_pred, target):
    """E|X-y| - 0.5 E|X-X'| (src/prob_unet_utils.py:237-268)."""
    return _native.ensemble_loss(ensemble_pred, target, kind="crps", alpha=0.0)


def wmse_ms_ssim_loss(pred, target, alpha=0.007, beta=0.048, lam=0.0, return_components=False, data_range=None):
    """lam*WMSE + (1-lam)*(1-MS-SSIM) (src/prob_unet_utils.py:270-305)."""
    if pred.dim() == 5:
        raise NotImplementedError("ensemble-mean input is not used by the reference elbo")
    res = _native.wmse_ms_ssim(pred, target, alpha, beta, lam, data_range)
    return res if return_components else res[0]

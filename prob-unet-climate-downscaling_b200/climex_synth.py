"""Synthetic ClimEx-shaped fields (pr, tasmin, tasmax) for tests and benchmarks.

There is no ClimEx NetCDF data on the box, so every measurement uses fields with the
same *structure* the reference's dataset produces (src/climex_utils.py:197-225):
inputs are piecewise-constant on s x s blocks (nearest-upsampled block means), targets
are the standardised residual ``hr - up`` (zero mean on every block).  Host-side
torch ops only -- data generation is not on the hot path.
"""
import torch
import torch.nn.functional as F


def make_fields(batch: int, height: int, width: int, lowres_scale: int = 16, seed: int = 1234):
    """Returns dict(inputs, targets, hr, lrinterp, std_hr) of fp32 NCHW CPU tensors.

    ``residual_to_hr(targets)`` reproduces ``hr``: hr == lrinterp + targets * std_hr.
    """
    g = torch.Generator().manual_seed(seed)
    hr = F.avg_pool2d(torch.randn(batch, 3, height + 8, width + 8, generator=g), 9, 1)
    hr = hr / hr.std()
    lr = F.avg_pool2d(hr, lowres_scale)
    up = F.interpolate(lr, scale_factor=lowres_scale)            # nearest, as the reference
    sigma = up.std()
    return {
        "inputs": (up / sigma).contiguous(),
        "targets": ((hr - up) / sigma).contiguous(),
        "hr": hr.contiguous(),
        "lrinterp": up.contiguous(),
        "std_hr": sigma.reshape(1, 1, 1, 1).expand(1, 3, 1, 1).contiguous(),
    }

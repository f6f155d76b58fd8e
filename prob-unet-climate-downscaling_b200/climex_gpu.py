"""Batched, GPU-side version of the reference dataset's transform (additive API).

``climex2torch.__getitem__`` (src/climex_utils.py:197-225, type "lrinterp_to_residuals") coarsens one field by
``lowres_scale`` x ``lowres_scale`` block means, upsamples it back (nearest), standardises with per-cell statistics
over time (``compute_stats``, :255-264) and forms the residual -- per sample, on the DataLoader's main thread.
``ClimexBatchTransform`` does the same for a whole batch in one kernel launch (``csrc/climex.cu``) on fields that
already live in HBM, returning the same dictionary keys with a leading batch dimension.

    tr = ClimexBatchTransform(lowres_scale=16)
    tr.compute_stats(hr_all)                      # hr_all [T,3,H,W] on the GPU (or a pinned host tensor)
    batch = tr(hr_all[idx])                       # {"inputs","targets","hr","lr","lrinterp"}
    hr_pred = tr.residual_to_hr(model(batch["inputs"], training=False), batch["lrinterp"])
"""
import ctypes as C

import torch

import _native as N


class ClimexBatchTransform:
    def __init__(self, lowres_scale=4, epsilon=1e-10):
        self.lowres_scale = int(lowres_scale)
        self.epsilon = float(epsilon)              # src/climex_utils.py:86
        self.lrstats = None                        # ((mean_lr, std_lr), (mean_hrdim, std_hrdim)) like the reference

    def compute_stats(self, hr):
        """hr [T,C,H,W] -> ((mean, std) [C,H/s,W/s], (mean_hrdim, std_hrdim) [C,H,W])  (src/climex_utils.py:255-264)."""
        hr = hr.cuda(non_blocking=True).contiguous().float()
        N.require_cuda(hr)
        T, Cc, H, W = hr.shape
        s = self.lowres_scale
        mean = torch.empty(Cc, H // s, W // s, device=hr.device, dtype=torch.float32)
        std = torch.empty_like(mean)
        N.check(N.lib().pub_climex_stats(N.ptr(hr), T, Cc, H, W, s, N.ptr(mean), N.ptr(std), N.stream()), "pub_climex_stats")
        up = lambda t: t.repeat_interleave(s, dim=1).repeat_interleave(s, dim=2)
        self.lrstats = ((mean, std), (up(mean), up(std)))
        return self.lrstats

    def __call__(self, hr):
        """hr [B,C,H,W] -> dict with the keys of climex2torch.__getitem__ (batched, on the device)."""
        if self.lrstats is None:
            raise RuntimeError("compute_stats(hr_all) must be called first (the reference computes it lazily from the whole dataset)")
        hr = hr.cuda(non_blocking=True).contiguous().float()
        B, Cc, H, W = hr.shape
        s = self.lowres_scale
        inputs, targets, lrinterp = torch.empty_like(hr), torch.empty_like(hr), torch.empty_like(hr)
        lr = torch.empty(B, Cc, H // s, W // s, device=hr.device, dtype=torch.float32)
        mean, std = self.lrstats[0]
        N.check(N.lib().pub_climex_transform(N.ptr(hr), N.ptr(mean), N.ptr(std), B, Cc, H, W, s, C.c_float(self.epsilon),
                                             N.ptr(inputs), N.ptr(targets), N.ptr(lrinterp), N.ptr(lr), N.stream()),
                "pub_climex_transform")
        return {"inputs": inputs, "targets": targets, "hr": hr, "lr": lr, "lrinterp": lrinterp}

    # src/climex_utils.py:277-285
    def invstand_residual(self, standardized_residual):
        return standardized_residual * (self.lrstats[1][1] + self.epsilon)

    def residual_to_hr(self, residual, lrinterp):
        return lrinterp + self.invstand_residual(residual)

"""GPU data transform (csrc/climex.cu, climex_gpu.ClimexBatchTransform) against fixtures from the reference's real
climex2torch class and against the oracle at a larger size."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import probunet_oracle as O

pytestmark = pytest.mark.gpu


def test_transform_matches_reference_fixtures():
    from climex_gpu import ClimexBatchTransform
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "climex_golden.npz"))
    hr = torch.from_numpy(g["hr"]).cuda()
    tr = ClimexBatchTransform(lowres_scale=int(g["scale"]))
    (mean, std), (mean_hr, std_hr) = tr.compute_stats(hr)
    # fp32 path tolerance 1e-4 (BASELINE.json): the temperature cell means are ~280 with a spread of ~0.5 over time, so
    # their std carries the f32 rounding of the means (3e-5 relative between two summation orders)
    assert rel_err(mean, g["mean_lr"]) < 1e-6 and rel_err(std, g["std_lr"]) < 1e-4
    assert rel_err(mean_hr, g["mean_hr"]) < 1e-6 and rel_err(std_hr, g["std_hr"]) < 1e-4
    batch = tr(hr[torch.from_numpy(g["idx"]).cuda()])
    # inputs = (cell mean - time mean) / std subtracts two f32 numbers near 280 K whose difference is ~0.5: one ulp of
    # either (3e-5) is 6e-5 of the result, before the std error -> 3e-4; the uncancelled outputs are at 1e-6
    for k, tol in (("inputs", 3e-4), ("lr", 1e-6), ("lrinterp", 1e-6)):
        assert rel_err(batch[k], g[k]) < tol, (k, rel_err(batch[k], g[k]))
    # targets are O(50) standardised residuals: they inherit the 3e-5 relative error of std
    assert rel_err(batch["targets"], g["targets"]) < 1e-4, rel_err(batch["targets"], g["targets"])
    res = torch.from_numpy(g["residual"]).cuda()
    assert rel_err(tr.residual_to_hr(res, batch["lrinterp"]), g["residual_to_hr"]) < 1e-6


def test_transform_full_size_round_trip():
    """B = 64 fields of 128 x 128, lowres_scale 16 (src/main.py:30): residual_to_hr(targets, lrinterp) gives hr back,
    inputs are constant on every 16 x 16 cell, targets have zero mean on every cell, and the oracle agrees."""
    from climex_gpu import ClimexBatchTransform
    g = torch.Generator().manual_seed(3)
    hr = torch.randn(96, 3, 128, 128, generator=g) * 4 + 2
    tr = ClimexBatchTransform(lowres_scale=16)
    tr.compute_stats(hr.cuda())
    b = tr(hr[:64].cuda())
    back = tr.residual_to_hr(b["targets"], b["lrinterp"])
    assert float((back - b["hr"]).abs().max()) < 1e-4
    cells = b["inputs"].reshape(64, 3, 8, 16, 8, 16)
    assert float((cells - cells[:, :, :, :1, :, :1]).abs().max()) == 0.0
    assert float(torch.nn.functional.avg_pool2d(b["targets"], 16).abs().max()) < 1e-5
    stats = O.climex_compute_stats(hr, 16)
    ref = O.climex_getitem(hr[:64], stats, 16)
    assert rel_err(b["inputs"], ref["inputs"]) < 1e-4 and rel_err(b["lrinterp"], ref["lrinterp"]) < 1e-6
    assert rel_err(b["targets"], ref["targets"]) < 1e-4

"""the ensemble post-processing diagnostics of results.ipynb on the GPU (csrc/spectra.cu) --
radially averaged PSD (cell 4: torch.fft.fftn + scipy binned_statistic per field on the host) and the value histograms
(cell 15: np.histogram) -- against fixtures produced by the notebook's own code and against the oracle at 128 x 128."""
import os

import numpy as np
import pytest
import torch

from oracle import probunet_oracle as O

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "climex_golden.npz"))


def test_radial_psd_matches_the_notebook_fixtures():
    import metrics
    import _native as N
    pf = torch.from_numpy(G["psd_fields"]).cuda()
    per, _ = N.radial_psd(pf, transfo=False, units=False)
    np.testing.assert_allclose(per[0, 1].cpu().numpy(), G["psd_single"], rtol=2e-4)      # psd() of one raw field
    for tf, key in ((True, "psd_tensor_transfo"), (False, "psd_tensor_plain")):
        out = metrics.compute_psd_tensor(pf, transfo=tf)
        got = np.stack([out[v] for v in ("pr", "tasmin", "tasmax")])
        np.testing.assert_allclose(got, G[key], rtol=2e-4)
        np.testing.assert_allclose(out["k"], G["psd_k"], rtol=0, atol=0)
    out5 = metrics.compute_psd_tensor(pf.reshape(2, 3, 3, 32, 32), transfo=True)          # [T, M, 3, H, W]
    np.testing.assert_allclose(np.stack([out5[v] for v in ("pr", "tasmin", "tasmax")]), G["psd_tensor_5d"], rtol=2e-4)


def test_radial_psd_at_128_matches_the_oracle_and_parseval():
    import _native as N
    g = torch.Generator().manual_seed(3)
    f = torch.randn(5, 3, 128, 128, generator=g) * torch.tensor([1.5, 4.0, 2.0]).view(1, 3, 1, 1) + torch.tensor([0.3, 280.0, 5.0]).view(1, 3, 1, 1)
    per, mean = N.radial_psd(f.cuda(), transfo=True, units=True)
    ref = O.compute_psd_tensor(f, True)
    np.testing.assert_allclose(mean.cpu().numpy(), ref, rtol=3e-4)
    assert per.shape == (5, 3, 64) and bool(torch.isfinite(per).all())
    # one plane wave: all the power sits in the bin of its wavenumber
    yy, xx = torch.meshgrid(torch.arange(128.0), torch.arange(128.0), indexing="ij")
    wave = torch.cos(2 * np.pi * (5 * xx + 12 * yy) / 128).reshape(1, 1, 128, 128)
    p1, _ = N.radial_psd(wave.cuda())
    k = int(round(float(np.hypot(5, 12)))) - 1                                          # bin [12.5, 13.5) holds |k| = 13
    assert int(p1[0, 0].argmax()) == k and float(p1[0, 0, k]) > 1e3 * float(p1[0, 0].sum() - p1[0, 0, k])
    with pytest.raises(N.NativeError):
        N.radial_psd(torch.zeros(1, 1, 96, 96, device="cuda"))


def test_histogram_matches_numpy():
    import metrics
    counts = metrics.value_histogram(torch.from_numpy(G["hist_values"]), G["hist_edges"])
    np.testing.assert_array_equal(counts, G["hist_counts"])
    g = torch.Generator().manual_seed(4)
    v = torch.randn(3_000_000, generator=g) * 7 - 2
    edges = np.linspace(float(v.min()), float(v.max()), 100)
    np.testing.assert_array_equal(metrics.value_histogram(v, edges), np.histogram(v.numpy(), bins=edges)[0])
    e2 = np.array([-1.0, 0.0, 0.5, 3.0])                                                 # uneven edges, values outside dropped
    np.testing.assert_array_equal(metrics.value_histogram(v, e2), np.histogram(v.numpy(), bins=e2)[0])

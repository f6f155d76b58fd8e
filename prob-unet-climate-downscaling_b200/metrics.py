"""Ensemble metrics -- host mirror of the reference's ``metrics.py`` (src/metrics.py:11-71).

Same function names, arguments and return structure; the per-(t,var) CRPS (pysteps'
Hersbach CRPS == E|X-y| - 0.5 E|X-X'|) and the MAE of the ensemble mean are computed by one
sm_100a kernel (csrc/latent_loss.cu) instead of T x 3 numpy calls.
"""
import torch

import _native as N

_VARS = ["pr", "tasmin", "tasmax"]


def _dev(x):
    x = torch.as_tensor(x)
    return x if x.is_cuda else x.cuda()


def crps_over_groundtruth(hr, preds):
    """hr (T,3,H,W), preds (T,M,3,H,W) in real units -> (means dict, per-timestep arrays dict)."""
    hr, preds = _dev(hr), _dev(preds)
    assert hr.shape == (preds.shape[0],) + tuple(preds.shape[2:])
    crps, _ = N.ensemble_metrics(preds, hr)
    c = crps.double().cpu().numpy()
    return ({v: float(c[:, i].mean()) for i, v in enumerate(_VARS)}, {v: c[:, i].copy() for i, v in enumerate(_VARS)})


def compute_mae(ground_truth, predictions):
    """MAE of the ensemble mean (or of a deterministic prediction) per variable."""
    gt, pr = _dev(ground_truth), _dev(predictions)
    if pr.dim() == 4:
        pr = pr.unsqueeze(1)
    _, mae = N.ensemble_metrics(pr, gt)
    m = mae.cpu().numpy()
    return ({v: float(m[:, i].mean()) for i, v in enumerate(_VARS)}, {v: m[:, i].copy() for i, v in enumerate(_VARS)})


def ensemble_scores_from_residuals(residual_preds, hr, lrinterp, std_hr):
    """Additive API: members are standardised residuals; residual_to_hr + inverse transforms
    (src/climex_utils.py:277-285, results.ipynb cell 2) are fused into the metric kernel."""
    return N.ensemble_metrics(_dev(residual_preds), _dev(hr), _dev(lrinterp), _dev(std_hr))


def compute_psd_tensor(data, transfo):
    """results.ipynb cell 4 (``compute_psd_tensor`` + ``psd``): radially averaged power spectral density of every
    (sample[, member], variable) field in real units, mean over samples -> {'pr': P(k), 'tasmin': P(k), 'tasmax': P(k)}
    (numpy, length H/2), plus ``kvals`` under the key 'k'.  data: [T,3,H,W] or [T,M,3,H,W]; ``transfo`` as in the
    notebook (True: data still in the stored-transform domain)."""
    d = _dev(data)
    if d.dim() == 5:
        d = d.reshape(-1, *d.shape[2:])
    _, mean = N.radial_psd(d, transfo=bool(transfo), units=True)
    m = mean.double().cpu().numpy()
    H = d.shape[-1]
    out = {v: m[i].copy() for i, v in enumerate(_VARS)}
    out["k"] = 0.5 * (torch.arange(0.5, H // 2 + 1, 1.0)[1:] + torch.arange(0.5, H // 2 + 1, 1.0)[:-1]).numpy()
    return out


def value_histogram(values, edges):
    """results.ipynb cell 15: ``np.histogram(values, bins=edges)`` counts (the caller takes log(count + 1))."""
    return N.histogram(_dev(values), edges).cpu().numpy()

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

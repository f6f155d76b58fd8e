"""Op-level parity of the GroupNorm (+FiLM) + SiLU (+dropout) (+2x resample) kernels (csrc/norm.cu, through
pub_groupnorm_silu_forward / _backward) against the ATen calls they replace: F.group_norm (src/networks.py:105-107),
silu / addcmul (:168-173), F.dropout (:177) and the box resample of the following Conv2d (:83-87), forward and
backward, at every (C, groups) pair of the canonical network incl. the non-power-of-two concats (96, 192, 384)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}


def _ref(x_nchw, gamma, beta, film, resample, keep=None, p=0.0):
    C = x_nchw.shape[1]
    u = F.group_norm(x_nchw, min(32, C // 4), gamma, beta, eps=1e-5)
    if film is not None:
        scale, shift = film.reshape(1, -1, 1, 1).chunk(2, dim=1)
        u = torch.addcmul(shift, u, scale + 1)
    y = F.silu(u)
    if keep is not None:
        y = y * keep / (1.0 - p)
    if resample == 1:
        y = F.avg_pool2d(y, 2)
    elif resample == 2:
        y = y.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return y


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,H,film,resample", [(32, 32, False, 0), (64, 16, True, 0), (96, 32, False, 0), (128, 16, True, 0),
                                               (192, 16, False, 0), (256, 16, True, 0), (384, 8, False, 0), (512, 16, False, 0),
                                               (64, 32, False, 1), (128, 16, False, 2), (8, 24, False, 0)])
def test_groupnorm_silu_forward_backward_match_aten(dtype, C, H, film, resample):
    import _native as N
    g = torch.Generator(device="cuda").manual_seed(C + H)
    B, W = 3, H + 8
    x = (torch.randn(B, H, W, C, device="cuda", generator=g) * 1.7 + 0.4).to(dtype)
    gamma = torch.randn(C, device="cuda", generator=g) * 0.3 + 1.0
    beta = torch.randn(C, device="cuda", generator=g) * 0.2
    fl = torch.randn(2 * C, device="cuda", generator=g) * 0.2 if film else None
    xs = x.clone().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in ([gamma, beta] + ([fl] if film else []))]
    y = N.groupnorm_silu_nhwc(xs, ps[0], ps[1], ps[2] if film else None, resample=resample)
    dy = torch.randn(y.shape, device="cuda", generator=g).to(dtype)
    y.backward(dy)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    pr = [t.clone().requires_grad_(True) for t in ([gamma, beta] + ([fl] if film else []))]
    yr = _ref(xr, pr[0], pr[1], pr[2] if film else None, resample)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    tol = TOL[dtype]
    assert rel_err(y.float().permute(0, 3, 1, 2), yr) < tol
    assert rel_err(xs.grad.float().permute(0, 3, 1, 2), xr.grad) < 2 * tol
    for a, b in zip(ps, pr):
        assert rel_err(a.grad, b.grad) < 2 * tol, (a.shape, rel_err(a.grad, b.grad))


def test_groupnorm_dropout_is_a_bernoulli_mask_applied_consistently_forward_and_backward():
    """F.dropout's generator is replaced by Philox4x32-10 keyed by (seed, subsequence, element): distributional
    parity only, so check the structure -- every output is either 0 or reference / (1 - p), ~90 % are kept, the same
    (seed, subseq) reproduces the mask, another subsequence does not, and the backward pass uses the forward's mask."""
    import _native as N
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W, C, p = 2, 32, 32, 64, 0.1
    x = torch.randn(B, H, W, C, device="cuda", generator=g)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    xs = x.clone().requires_grad_(True)
    y = N.groupnorm_silu_nhwc(xs, gamma, beta, None, p_drop=p, seed=77, subseq=3)
    ref = _ref(x.permute(0, 3, 1, 2), gamma, beta, None, 0).permute(0, 2, 3, 1)
    keep = (y != 0) | (ref == 0)
    assert abs(float(keep.float().mean()) - (1 - p)) < 5e-3
    assert rel_err(y[keep], ref[keep] / (1 - p)) < 1e-4
    y2 = N.groupnorm_silu_nhwc(x, gamma, beta, None, p_drop=p, seed=77, subseq=3)
    y3 = N.groupnorm_silu_nhwc(x, gamma, beta, None, p_drop=p, seed=77, subseq=4)
    assert torch.equal(y, y2) and not torch.equal(y, y3)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = _ref(xr, gamma, beta, None, 0, keep=keep.permute(0, 3, 1, 2).float(), p=p)
    yr.backward(dy.permute(0, 3, 1, 2))
    assert rel_err(xs.grad.permute(0, 3, 1, 2), xr.grad) < 2e-4

"""Module- and model-level parity of the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors of the real reference.  Tolerances from BASELINE.json: fp32 path rel-err <= 1e-4 on outputs and
losses, bf16 path <= 1e-2, CRPS within 0.5 %."""
import numpy as np
import pytest
import torch

from helpers import canonical_model, grad_projection, proj_close, rel_err, unpack_masks
from oracle import probunet_oracle as O

pytestmark = pytest.mark.gpu
CFG = O.ProbUNetCfg()
TOL = {"fp32": 1e-4, "bf16": 1e-2}      # BASELINE.json north_star: outputs, losses, KL
# Gradients are not covered by north_star's output tolerances; every one of the 391 tensors is compared ELEMENT-WISE
# (||g - g_ref|| / ||g_ref||) with the live oracle, whose own gradients are pinned by signed projections of the real
# reference's (tests/test_oracle_golden.py).  bf16 activations put ~1e-2 of rounding noise on every backward operand.
GTOL = {"fp32": 2e-3, "bf16": 1.2e-1}
GTOL_NORM = {"fp32": 2e-3, "bf16": 6e-2}


def oracle_grads(sd, fn):
    """Runs `fn(full_state_dict)` through the oracle with every parameter a leaf; -> (outputs, {name: grad})."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    out = fn(full)
    out[0].backward()
    return out, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


@pytest.fixture(scope="module", params=["fp32", "bf16"])
def setup(request):
    m = canonical_model(compute_dtype=request.param, device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    return request.param, m, sd


def _inputs(golden, dev="cuda"):
    return tuple(torch.from_numpy(golden[k]).to(dev) for k in ("A_x", "A_y", "A_eps"))


def test_unet_features(setup, golden):
    name, m, sd = setup
    x, _, _ = _inputs(golden)
    with torch.no_grad():
        f = m.unet(x)
    assert f.shape == (2, 32, 64, 64) and f.dtype == torch.float32
    assert rel_err(f, golden["A_unet"]) < TOL[name]


def test_prior_posterior_heads_and_kl(setup, golden):
    import _native as N
    name, m, sd = setup
    x, y, _ = _inputs(golden)
    with torch.no_grad():
        p, q = m.prior(x), m.posterior(x, y)
        assert rel_err(p.base_dist.loc, golden["A_prior_mu"]) < TOL[name]
        assert rel_err(p.base_dist.scale, golden["A_prior_sigma"]) < TOL[name]
        assert rel_err(q.base_dist.loc, golden["A_post_mu"]) < TOL[name]
        assert rel_err(q.base_dist.scale, golden["A_post_sigma"]) < TOL[name]
        kl = N.kl_normal(q.base_dist.loc, q.base_dist.scale, p.base_dist.loc, p.base_dist.scale)
        assert rel_err(kl, golden["A_kl"]) < TOL[name]
        # torch.distributions interop (latent-exploration scripts use kl_divergence / base_dist)
        kl_t = torch.distributions.kl.kl_divergence(q, p)
        assert rel_err(kl, kl_t) < 1e-5


def test_lazy_elbo_scalars_match_the_synchronous_ones(golden):
    """sync_scalars = "lazy": the reconstruction term comes back as a LazyScalar whose device -> pinned-host copy was
    enqueued inside elbo(); same value as the .item() float, gradients untouched."""
    from lazy_scalar import LazyScalar
    x, y, eps = _inputs(golden)
    m = canonical_model(compute_dtype="fp32", device="cuda")
    out = {}
    for mode in (True, "lazy", False):
        m.sync_scalars = mode
        m.zero_grad(set_to_none=True)
        total, recon, kl = m.elbo(x, y, None, M=eps.shape[0], eps=eps)
        total.backward()
        out[mode] = (float(total), recon[0], m.fcomb.layers[4].bias.grad.clone())
    assert isinstance(out[True][1], float) and isinstance(out["lazy"][1], LazyScalar) and torch.is_tensor(out[False][1])
    assert float(out["lazy"][1]) == out[True][1] == float(out[False][1])
    assert out["lazy"][0] == out[True][0] and torch.equal(out["lazy"][2], out[True][2])


def test_encoder_precision_policy(golden, monkeypatch):
    """The Gaussian encoders of a bf16 model run in tf32 (f32 storage).  Measured against the golden mu / sigma of the
    real reference: all-bf16 encoders miss the 1e-2 bar (1.3e-2); tf32 behind a bf16 full-resolution stage
    (tf32_bf16s0, opt-in) stays at 2.2e-3 -- 2.5x tf32's 0.9e-3 -- and is 1 % faster per step, but its bf16 gradient
    tensors push the element-wise error of the prior's first-stage weight gradients to 0.17 (bar 0.12), so it is not
    the default."""
    import _native as N
    x, y, _ = _inputs(golden)
    err = {}
    for enc in ("bf16", "tf32_bf16s0", "tf32"):
        monkeypatch.setenv("PROBUNET_B200_ENCODER_DTYPE", enc)
        m = canonical_model(compute_dtype="bf16", device="cuda")
        with torch.no_grad():
            p, q = m.prior(x), m.posterior(x, y)
        assert m.prior.engine().dtype == {"bf16": N.BF16, "tf32_bf16s0": N.TF32_BF16S0, "tf32": N.TF32}[enc]
        err[enc] = max(rel_err(p.base_dist.scale, golden["A_prior_sigma"]), rel_err(q.base_dist.scale, golden["A_post_sigma"]),
                       rel_err(p.base_dist.loc, golden["A_prior_mu"]), rel_err(q.base_dist.loc, golden["A_post_mu"]))
    monkeypatch.delenv("PROBUNET_B200_ENCODER_DTYPE")
    assert N.resolve_encoder_dtype("bf16") == N.TF32 and N.resolve_encoder_dtype("fp32") == N.F32
    assert err["tf32"] < 2e-3, err
    assert err["tf32_bf16s0"] < 4e-3, err
    assert err["tf32_bf16s0"] < 0.4 * err["bf16"], err


def test_forward_training_and_prior_paths(setup, golden):
    name, m, sd = setup
    x, y, eps = _inputs(golden)
    with torch.no_grad():
        o1 = m(x, y, training=True, eps=eps[0])
        o2 = m(x, None, t=None, training=False, eps=eps[1])
    assert rel_err(o1, golden["A_fwd_train"]) < TOL[name]
    assert rel_err(o2, golden["A_fwd_prior"]) < TOL[name]
    assert m.posterior_latent_space is not None and m.prior_latent_space is not None


def test_fcomb_public_call_with_expanded_features(setup, golden):
    name, m, sd = setup
    x, y, eps = _inputs(golden)
    with torch.no_grad():
        feat = m.unet(x)
        q = m.posterior(x, y)
        z = q.base_dist.loc + q.base_dist.scale * eps[2]
        out = m.fcomb(feat, z)
        assert rel_err(out, golden["A_fcomb"]) < TOL[name]
        # non-contiguous expand() view, as src/latent_exploration_posterior.py:122-123 does
        fe = feat[:1].expand(2, -1, -1, -1)
        oe = m.fcomb(fe, z)
        ref = O.fcomb(sd, feat[:1].cpu().expand(2, -1, -1, -1), z.cpu())
        assert rel_err(oe, ref) < 1e-4


def _check_grads(m, names, ref_norms, tol, ref_grads=None, elem_tol=None, ref_proj=None):
    """Every gradient against (a) the golden norm of the real reference, (b) ELEMENT-WISE the oracle's gradient,
    (c) the golden signed projections (catch a transposed / mirrored / permuted gradient with the right norm)."""
    bad, worst = [], (0.0, None)
    params = dict(m.named_parameters())
    for i, (n, r) in enumerate(zip(names, ref_norms)):
        n = str(n)
        g = params[n].grad
        assert g is not None, f"{n} has no gradient"
        gn = float(g.double().norm())
        if abs(gn - r) > tol * max(r, 1e-6) + 1e-9:
            bad.append(("norm", n, gn, float(r)))
        if ref_grads is not None and float(r) > 1e-7:
            e = rel_err(g, ref_grads[n])
            worst = max(worst, (e, n))
            if e > elem_tol:
                bad.append(("elem", n, e))
        if ref_proj is not None and float(r) > 1e-7:
            if not proj_close(grad_projection(g), ref_proj[i], float(r), 3 * (elem_tol or tol)):
                bad.append(("proj", n, grad_projection(g), list(ref_proj[i]), float(r)))
    assert not bad, f"{len(bad)} gradient checks failed (worst element-wise {worst}): {bad[:8]}"
    return worst


def test_elbo_afcrps_loss_and_all_gradients(setup, golden):
    name, m, sd = setup
    x, y, eps = _inputs(golden)
    m.loss_type = "afcrps"
    m.zero_grad(set_to_none=True)
    total, recon, kl = m.elbo(x, y, None, M=3, eps=eps)
    total.backward()
    assert abs(float(total) - float(golden["A_afcrps_total"])) / abs(float(golden["A_afcrps_total"])) < TOL[name]
    assert abs(recon[0] - float(golden["A_afcrps_crps"])) / float(golden["A_afcrps_crps"]) < 5e-3   # CRPS within 0.5 %
    assert rel_err(kl, golden["A_kl"]) < TOL[name]
    _, ref = oracle_grads(sd, lambda s_: O.elbo(s_, CFG, x.cpu(), y.cpu(), eps.cpu(), "afcrps"))
    _check_grads(m, list(golden["grad_names"]), golden["A_afcrps_gradnorm"], GTOL_NORM[name], ref, GTOL[name],
                 golden["A_afcrps_gradproj"])
    for k in golden.files:
        if k.startswith("A_afcrps_grad::"):
            g = dict(m.named_parameters())[k.split("::")[1]].grad
            assert rel_err(g, golden[k]) < GTOL[name], (k, rel_err(g, golden[k]))
    # parameters whose input is identically zero get exact zero gradients, not None
    assert float(m.unet.map_label.weight.grad.abs().sum()) == 0.0


def test_elbo_l1_loss_and_gradients(setup, golden):
    name, m, sd = setup
    x, y, eps = _inputs(golden)
    m.loss_type = "l1"
    m.zero_grad(set_to_none=True)
    total, per_var, kl, kl2 = m.elbo(x, y, None, eps=eps[:1])
    total.backward()
    m.loss_type = "afcrps"
    assert abs(float(total) - float(golden["A_l1_total"])) / abs(float(golden["A_l1_total"])) < TOL[name]
    assert len(per_var) == 3 and kl2.shape == (2,)
    _, ref = oracle_grads(sd, lambda s_: O.elbo(s_, CFG, x.cpu(), y.cpu(), eps[:1].cpu(), "l1"))
    _check_grads(m, list(golden["grad_names"]), golden["A_l1_gradnorm"], GTOL_NORM[name], ref, GTOL[name],
                 golden["A_l1_gradproj"])


def test_train_mode_dropout_matches_oracle_with_exported_masks(setup, golden):
    """Dropout uses the engine's own Philox stream (distributional parity with F.dropout); the masks are exported
    through pub_unet_dropout_mask and injected into the oracle -> exact parity of the train-mode forward."""
    name, m, sd = setup
    x, y, eps = _inputs(golden)
    eng = m.unet.engine()
    m.unet.train()
    try:
        with torch.no_grad():
            f = eng.forward(x, True, seed=1234)
        keys = [k for k in eng.block_keys if not k.endswith("_conv")]
        enc, dec = O.unet_topology(CFG.unet())
        assert keys == [b.key for b in enc + dec if not b.is_conv]
        masks = {k: eng.dropout_mask(k, 2, 64, 64, 1234).cpu() for k in keys}
        frac = np.mean([float(v.float().mean()) for v in masks.values()])
        assert abs(frac - 0.9) < 0.01                      # keep probability 1 - p
        with torch.no_grad():
            ref = O.unet_forward(sd, x.cpu(), CFG.unet(), drop_masks=masks)
        assert rel_err(f, ref) < TOL[name]
    finally:
        m.unet.eval()


def test_sample_members_and_metrics(setup, golden):
    import _native as N
    import metrics
    name, m, sd = setup
    x, y, _ = _inputs(golden)
    n = 6
    eps = torch.randn(n, 2, 32, generator=torch.Generator().manual_seed(9)).cuda()
    ens = m.sample(x, n, eps=eps)                                   # [B,n,3,H,W]
    assert ens.shape == (2, n, 3, 64, 64)
    with torch.no_grad():
        feat = O.unet_forward(sd, x.cpu(), CFG.unet())
        mu, sig = O.gaussian_encoder(sd, "prior", x.cpu(), None, CFG.num_filters)
        ref = torch.stack([O.fcomb(sd, feat, mu + sig * eps[i].cpu()) for i in range(n)], dim=1)
    assert rel_err(ens, ref) < TOL[name]
    # metrics kernel vs the oracle restatement of src/metrics.py on the same members
    hr = y.cpu() * 1.7 + 0.3
    crps_ref = O.crps_over_groundtruth(hr, ref)
    mae_ref = O.compute_mae(hr, ref)
    means, arrays = metrics.crps_over_groundtruth(hr, ref)
    assert abs(means["pr"] - crps_ref[:, 0].mean()) < 1e-5 * abs(crps_ref[:, 0].mean()) + 1e-7
    np.testing.assert_allclose(arrays["tasmax"], crps_ref[:, 2], rtol=2e-5)
    mm, ma = metrics.compute_mae(hr, ref)
    np.testing.assert_allclose(ma["tasmin"], mae_ref[:, 1], rtol=2e-5)
    # fused residual_to_hr + inverse transforms
    lr, std = torch.randn(2, 3, 64, 64) + 275.0, torch.tensor([1.5, 2.0, 2.5]).reshape(1, 3, 1, 1)
    real = O.residual_to_real(ref, lr.unsqueeze(1), std.unsqueeze(1))
    hr_real = O.residual_to_real(y.cpu(), lr, std)
    c2, a2 = metrics.ensemble_scores_from_residuals(ref, hr_real, lr, std)
    np.testing.assert_allclose(c2.cpu().numpy(), O.crps_over_groundtruth(hr_real, real), rtol=2e-3, atol=1e-4)
    np.testing.assert_allclose(a2.cpu().numpy(), O.compute_mae(hr_real, real), rtol=2e-3, atol=1e-4)


def test_pixel_targeted_sampling_equals_the_full_field_members(setup, golden):
    """SURVEY.md 8f rank 4 (test_return_levels.ipynb cell 2 reads one pixel of a full forward pass per member and day):
    sample_pixels evaluates fcomb at the requested pixels only; it must equal the full-field members there, and the
    oracle's full forward."""
    name, m, sd = setup
    x, _, _ = _inputs(golden)
    n = 5
    eps = torch.randn(n, 2, 32, generator=torch.Generator().manual_seed(19)).cuda()
    pix = [(0, 0), (63, 63), (17, 40), (5, 62), (33, 1), (8, 8), (60, 3)]
    vals = m.sample_pixels(x, n, pix, eps=eps)                      # [B,n,3,P]
    assert vals.shape == (2, n, 3, len(pix))
    full = m.sample(x, n, eps=eps)
    ys, xs = [p[0] for p in pix], [p[1] for p in pix]
    assert rel_err(vals, full[..., ys, xs]) < 1e-6 if name == "fp32" else rel_err(vals, full[..., ys, xs]) < 2e-5
    with torch.no_grad():
        feat = O.unet_forward(sd, x.cpu(), CFG.unet())
        mu, sig = O.gaussian_encoder(sd, "prior", x.cpu(), None, CFG.num_filters)
        ref = torch.stack([O.fcomb(sd, feat, mu + sig * eps[i].cpu()) for i in range(n)], dim=1)[..., ys, xs]
    assert rel_err(vals, ref) < TOL[name]
    # the per-pixel series in real units (return-level notebook formulas; c = 0 for tasmax there)
    hr_pix = 270.0 + 3.0 * vals.permute(0, 1, 3, 2).cpu()             # stand-in for residual_to_hr at the pixels
    for var in ("pr", "tasmin", "tasmax"):
        assert torch.isfinite(O.return_level_pixel_series(hr_pix, var)).all()
    with pytest.raises(IndexError):
        m.sample_pixels(x, 2, [(64, 0)])


def test_loss_kernels_known_answers(golden):
    import prob_unet_utils as U
    e, t = torch.from_numpy(golden["L_ens"]).cuda().requires_grad_(True), torch.from_numpy(golden["L_tgt"]).cuda()
    a = U.afcrps_loss(e, t, alpha=0.95)
    c = U.crps_loss(e, t)
    assert abs(float(a) - float(golden["L_afcrps"])) < 2e-6 and abs(float(c) - float(golden["L_crps"])) < 2e-6
    (a * 3.0).backward()
    ec = torch.from_numpy(golden["L_ens"]).requires_grad_(True)
    (O.afcrps_loss(ec, t.cpu()) * 3.0).backward()
    assert rel_err(e.grad, ec.grad) < 1e-5


def test_fused_adamw_matches_torch():
    from optim import FusedAdamW
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in [(7,), (33, 5), (4, 3, 3, 3), (1000,)]]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    o1 = FusedAdamW(ps, lr=1e-3, weight_decay=1e-2)
    o2 = torch.optim.AdamW(qs, lr=1e-3, weight_decay=1e-2)
    for step in range(5):
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, generator=g).cuda()
            p.grad, q.grad = gr.clone(), gr.clone()
        o1.step(); o2.step()
    for p, q in zip(ps, qs):
        assert rel_err(p, q) < 1e-6


def test_three_training_steps_track_the_oracle(golden):
    """3 AdamW steps (L1 ELBO, fp32 path) compared parameter-wise with the oracle + torch.optim.AdamW on CPU."""
    from optim import FusedAdamW
    m = canonical_model(compute_dtype="fp32", loss_type="l1", device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    x, y, eps = _inputs(golden)
    opt = FusedAdamW(m.parameters(), lr=1e-4)
    ref_opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4)
    for step in range(3):
        opt.zero_grad()
        total, *_ = m.elbo(x, y, None, eps=eps[step:step + 1])
        total.backward()
        opt.step()
        ref_opt.zero_grad()
        full = dict(sd); full.update(leaves)
        rt = O.elbo(full, CFG, x.cpu(), y.cpu(), eps[step:step + 1].cpu(), "l1")[0]
        rt.backward()
        for k, v in leaves.items():          # parameters with exactly-zero gradients still get weight decay
            if v.grad is None:
                v.grad = torch.zeros_like(v)
        ref_opt.step()
        assert abs(float(total) - float(rt)) / abs(float(rt)) < 2e-4, (step, float(total), float(rt))
    # Adam divides by sqrt(v): elements whose gradient is ~1e-8 turn 1e-6-level noise into O(lr) parameter
    # differences, so the per-tensor parameter tolerance is looser than the loss tolerance
    worst = max(rel_err(p, leaves[n]) for n, p in m.named_parameters())
    assert worst < 5e-3, worst


def test_wmse_msssim_kernel_value_and_gradient_match_the_oracle():
    """csrc/msssim.cu vs oracle.wmse_ms_ssim_loss (restatement of src/prob_unet_utils.py:270-305 + pytorch_msssim 1.0.0)
    and its autograd gradient, lam = 0 (the reference default: pure 1 - MS-SSIM) and lam = 0.3."""
    import prob_unet_utils as U
    g = torch.Generator().manual_seed(5)
    y = torch.nn.functional.avg_pool2d(torch.randn(2, 3, 136, 136, generator=g), 9, 1)
    y = y / y.std()
    for lam in (0.0, 0.3):
        x = (y + 0.3 * torch.randn(2, 3, 128, 128, generator=g)).requires_grad_(True)
        ref = O.wmse_ms_ssim_loss(x, y, lam=lam)
        ref[0].backward()
        xc = x.detach().cuda().requires_grad_(True)
        out = U.wmse_ms_ssim_loss(xc, y.cuda(), lam=lam, return_components=True)
        (out[0] * 2.0).backward()
        for a, b in zip(out, ref):
            assert abs(float(a) - float(b)) < 1e-5 * abs(float(b)) + 1e-7, (lam, float(a), float(b))
        assert rel_err(xc.grad, 2.0 * x.grad) < 1e-4, (lam, rel_err(xc.grad, 2.0 * x.grad))


def test_elbo_msssim_variant_at_128(golden):
    """The reference's ACTIVE elbo (src/prob_unet.py:229-267): 5-tuple return, golden values from the real reference
    (with the restated ms_ssim stubbed in, tests/golden/make_golden.py) at 128x128, B = 1.

    The de-zeroed seed-42 model predicts values of +-1900 against a target range of 18, so sigma_x^2 = E[x^2] - mu^2
    cancels ~7 digits in fp32: the ORACLE's own 1 - MS-SSIM moves by 0.85 % when its input moves by 2.6e-6 (measured,
    tools/msssim_conditioning_probe.py).  Tight parity of the kernel is therefore asserted on well-conditioned fields in
    test_wmse_msssim_kernel_value_and_gradient_match_the_oracle; here the ill-conditioned term gets a 2 % band and
    everything around it (total, KL, WMSE, arity, KL-driven gradients) the usual tolerance."""
    x, y, eps = (torch.from_numpy(golden[k]).cuda() for k in ("B_x", "B_y", "B_eps"))
    for name in ("fp32", "bf16"):
        m = canonical_model(compute_dtype=name, loss_type="mse+ssim", device="cuda")
        m.zero_grad(set_to_none=True)
        total, recon, kl, wmse, ms = m.elbo(x, y, None, M=eps.shape[0], eps=eps)
        total.backward()
        assert isinstance(recon, list) and isinstance(wmse, float) and isinstance(ms, float)
        assert abs(float(total) - float(golden["B_total"])) / abs(float(golden["B_total"])) < TOL[name]
        assert rel_err(kl, golden["B_kl"]) < TOL[name]
        assert abs(wmse - float(golden["B_wmse"])) / float(golden["B_wmse"]) < 2 * TOL[name]   # squares the output error
        assert abs(recon[0] - float(golden["B_recon"])) / abs(float(golden["B_recon"])) < 2e-2
        assert abs(ms - float(golden["B_msssim_loss"])) / float(golden["B_msssim_loss"]) < 2e-2
        names, norms = list(golden["grad_names"]), golden["B_gradnorm"]
        kl_driven = [i for i, n in enumerate(names) if n.startswith("prior.")]
        _check_grads(m, [names[i] for i in kl_driven], [norms[i] for i in kl_driven], GTOL_NORM[name])
        for n, p_ in m.named_parameters():
            assert p_.grad is not None and bool(torch.isfinite(p_.grad).all()), n


def test_dispatcher_ops_match_the_module_path(setup, golden):
    """torch.ops.probunet_b200.* (custom_ops.py) give the same numbers and gradients as the module path."""
    import custom_ops  # noqa: F401
    import _native as N
    import torch.nn.functional as F
    name, m, sd = setup
    ops = torch.ops.probunet_b200
    g = torch.Generator(device="cuda").manual_seed(4)
    # conv (+ReLU) forward/backward vs F.conv2d
    x = torch.randn(2, 32, 32, 32, device="cuda", generator=g).bfloat16().requires_grad_(True)
    w = (torch.randn(64, 32, 3, 3, device="cuda", generator=g) / 17).requires_grad_(True)
    b = torch.randn(64, device="cuda", generator=g).requires_grad_(True)
    y = ops.conv2d_nhwc(x, w, b, True)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr = x.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = w.detach().bfloat16().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, br, padding=1))
    yr.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_err(y.float().permute(0, 3, 1, 2), yr) < 1e-2
    assert rel_err(x.grad.float().permute(0, 3, 1, 2), xr.grad) < 2e-2 and rel_err(w.grad, wr.grad) < 2e-2
    assert rel_err(b.grad, br.grad) < 2e-2
    # fcomb op == module call, including gradients wrt features and z
    x_in, y_in, eps = _inputs(golden)
    with torch.no_grad():
        feat = m.unet(x_in)
    z = torch.randn(3, 2, 32, device="cuda", generator=g)
    params = [p for l in (m.fcomb.layers[0], m.fcomb.layers[2], m.fcomb.layers[4]) for p in (l.weight, l.bias)]
    f1, z1 = feat.clone().requires_grad_(True), z.clone().requires_grad_(True)
    o1 = ops.fcomb(f1, z1, *params)
    f2, z2 = feat.clone().requires_grad_(True), z.clone().requires_grad_(True)
    o2 = m.fcomb.forward_members(f2, z2)
    assert rel_err(o1, o2) < 1e-6
    go = torch.randn_like(o1)
    g1 = torch.autograd.grad(o1, [f1, z1], go)
    g2 = torch.autograd.grad(o2, [f2, z2], go)
    assert rel_err(g1[0], g2[0]) < 1e-5 and rel_err(g1[1], g2[1]) < 1e-5
    # losses + KL
    ens = o2.detach().clone().requires_grad_(True)
    l_op, _ = ops.ensemble_loss(ens, y_in, 0, 0.95)
    l_op.backward()
    ens2 = o2.detach().clone().requires_grad_(True)
    l_mod = N.ensemble_loss(ens2, y_in, "afcrps", 0.95)
    l_mod.backward()
    assert abs(float(l_op) - float(l_mod)) < 1e-6 * abs(float(l_mod)) and rel_err(ens.grad, ens2.grad) < 1e-6
    mq, sq, mp, sp = (torch.randn(4, 32, device="cuda", generator=g) for _ in range(4))
    sq, sp = sq.abs() + 0.1, sp.abs() + 0.1
    assert rel_err(ops.kl_normal(mq, sq, mp, sp), N.kl_normal(mq, sq, mp, sp)) < 1e-6


def test_fcomb_forward_tensor_core_kernel_close_to_f32_kernel(golden):
    """bf16 mode runs Fcomb.forward on tensor cores with every f32 operand split into two bf16 terms (hi + lo, three
    MMAs per product): it must reproduce the f32-FMA kernel on the same bf16 features to ~1e-5 (plain bf16 operands
    differ by 3.3e-3 and moved the afCRPS by 0.8 %), and stay inside the bf16 budget of the oracle."""
    import _native as N
    m = canonical_model(compute_dtype="bf16", device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    x, y, _ = _inputs(golden)
    z = torch.randn(7, 2, 32, generator=torch.Generator().manual_seed(21)).cuda()
    with torch.no_grad():
        feat = m.unet(x, _nhwc_out=True)
        out = {}
        try:
            for opt in (0, 1):
                N.lib().pub_debug_option(b"fcomb_fwd_mma", opt)
                out[opt] = N.fcomb_apply(m.fcomb, feat, z, nhwc=True)
                torch.cuda.synchronize()
        finally:
            N.lib().pub_debug_option(b"fcomb_fwd_mma", 1)
        ref_feat = O.unet_forward(sd, x.cpu(), CFG.unet())
        ref = torch.stack([O.fcomb(sd, ref_feat, z[i].cpu()) for i in range(7)], dim=1)
    e_kernels = rel_err(out[1], out[0])
    assert e_kernels < 5e-5, e_kernels
    assert rel_err(out[1], ref) < TOL["bf16"], (rel_err(out[1], ref), rel_err(out[0], ref))


def test_fcomb_forward_tf32_kernel_for_large_ensembles(golden):
    """Ensemble sampling (M > 32) runs Fcomb.forward on tf32 MMAs (one MMA per product instead of the three of the
    bf16 hi + lo split): within 1e-3 of the f32-FMA kernel on the same features, CRPS / MAE of a 100-member ensemble
    within 0.5 % (BASELINE.json), and the M > 32 dispatch really takes it."""
    import _native as N
    m = canonical_model(compute_dtype="bf16", device="cuda")
    x, y, _ = _inputs(golden)
    z = torch.randn(100, 2, 32, generator=torch.Generator().manual_seed(23)).cuda()
    with torch.no_grad():
        feat = m.unet(x, _nhwc_out=True)
        out = {}
        try:
            for opt in (0, 2, 1):
                N.lib().pub_debug_option(b"fcomb_fwd_mma", opt)
                out[opt] = N.fcomb_apply(m.fcomb, feat, z, nhwc=True)
                torch.cuda.synchronize()
        finally:
            N.lib().pub_debug_option(b"fcomb_fwd_mma", 1)
    assert rel_err(out[2], out[0]) < 1e-3, rel_err(out[2], out[0])
    assert torch.equal(out[1], out[2])                                  # M = 100 > 32: the default IS the tf32 kernel
    hr = y * 1.7 + 0.3
    c0, a0 = N.ensemble_metrics(out[0], hr)
    c2, a2 = N.ensemble_metrics(out[2], hr)
    assert float(((c2 - c0).abs() / c0.abs()).max()) < 5e-3 and float(((a2 - a0).abs() / a0.abs()).max()) < 5e-3


def test_deterministic_unet_config2_forward_backward():
    """BASELINE configs[1]: networks.UNet(img_resolution, in_channels=3, out_channels=3, label_dim=0) as built by
    src/deterministic_unet_main.py:52 (model_channels 16, channel_mult [1,4,8,16]) trained with MSE
    (src/trainmodel.py:158-160).  16- and 3-channel layers take the fp32-FMA conv path, the rest tcgen05."""
    import networks
    from helpers import dezero
    torch.manual_seed(42)
    net = networks.UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False)
    dezero(net)
    sd = {"unet." + k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.UNetCfg(in_channels=3, out_channels=3, model_channels=16, channel_mult=(1, 4, 8, 16), label_dim=0,
                    img_resolution=(64, 64))
    g = torch.Generator().manual_seed(6)
    x, y = torch.randn(2, 3, 64, 64, generator=g), torch.randn(2, 3, 64, 64, generator=g)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    ref = O.unet_forward(full, x, cfg)
    torch.nn.functional.mse_loss(ref, y).backward()
    # per-tensor gradient rel-err (stricter than the norm comparison of the Prob U-Net tests): bf16 rounding of the
    # activations shows up as ~8 % on the smallest GroupNorm / FiLM gradients
    for name, tol, gtol in (("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 1.2e-1)):
        net.compute_dtype = name
        m = net.cuda().eval()
        m.zero_grad(set_to_none=True)
        out = m(x.cuda(), class_labels=None)            # the keyword train_step passes (src/trainmodel.py:158)
        assert out.shape == (2, 3, 64, 64)
        assert rel_err(out, ref) < tol, (name, rel_err(out, ref))
        torch.nn.functional.mse_loss(out, y.cuda()).backward()
        bad = []
        for k, p in m.named_parameters():
            r = leaves["unet." + k].grad
            if r is None:                                # map_label is absent / affine.weight sees a zero input
                assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k
                continue
            if float(r.norm()) > 1e-7 and rel_err(p.grad, r) > gtol:
                bad.append((k, rel_err(p.grad, r)))
        assert not bad, (name, bad[:6])


def test_posterior_latent_sweep_shapes_at_256():
    """BASELINE configs[4] (src/latent_exploration_posterior.py): 256 x 256 grids, posterior means, one U-Net pass and
    a grid of fcomb decodes on an expanded (non-contiguous) feature view -- against the oracle."""
    m = canonical_model(latent_dim=8, compute_dtype="fp32", device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    cfg = O.ProbUNetCfg(latent_dim=8)
    g = torch.Generator().manual_seed(10)
    x, y = torch.randn(1, 3, 256, 256, generator=g), torch.randn(1, 3, 256, 256, generator=g)
    with torch.no_grad():
        q = m.posterior(x.cuda(), y.cuda())
        feat = m.unet(x.cuda())
        zs = q.base_dist.loc + q.base_dist.scale * torch.linspace(-3, 3, 4, device="cuda").unsqueeze(1)   # [4, L]
        dec = m.fcomb(feat.expand(4, -1, -1, -1), zs)
        mu_r, sig_r = O.gaussian_encoder(sd, "posterior", x, y, cfg.num_filters)
        feat_r = O.unet_forward(sd, x, cfg.unet())
        dec_r = O.fcomb(sd, feat_r.expand(4, -1, -1, -1), zs.cpu())
    assert rel_err(q.base_dist.loc, mu_r) < 1e-4 and rel_err(q.base_dist.scale, sig_r) < 1e-4
    assert feat.shape == (1, 32, 256, 256) and rel_err(feat, feat_r) < 1e-4
    assert dec.shape == (4, 3, 256, 256) and rel_err(dec, dec_r) < 1e-4


@pytest.mark.parametrize("name", ["fp32", "bf16"])
def test_elbo_on_a_non_square_grid_matches_the_oracle(name):
    """96 x 64 fields (not a power of two, H != W): the pyramid is 96x64 / 48x32 / 24x16 / 12x8, so the halo conv
    (H % 16 == 0), the per-tap conv and the FMA fallback (24x16, 12x8) are all on the path; loss and every
    gradient norm against the oracle."""
    m = canonical_model(compute_dtype=name, device="cuda")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    g = torch.Generator().manual_seed(21)
    x, y = torch.randn(2, 3, 96, 64, generator=g), torch.randn(2, 3, 96, 64, generator=g)
    eps = torch.randn(3, 2, CFG.latent_dim, generator=g)
    rt = O.elbo(full, CFG, x, y, eps, "afcrps")
    rt[0].backward()
    m.loss_type = "afcrps"
    m.zero_grad(set_to_none=True)
    total, recon, kl = m.elbo(x.cuda(), y.cuda(), None, M=3, eps=eps.cuda())
    total.backward()
    assert abs(float(total) - float(rt[0])) / abs(float(rt[0])) < TOL[name], (float(total), float(rt[0]))
    bad = []
    for n, p in m.named_parameters():
        r = leaves[n].grad
        if r is None or float(r.norm()) < 1e-7:
            continue
        e = rel_err(p.grad, r)                      # element-wise, every tensor
        if e > GTOL[name]:
            bad.append((n, e))
    assert not bad, bad[:6]

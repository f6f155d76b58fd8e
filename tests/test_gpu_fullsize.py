"""Size-independent properties at BASELINE.json's FULL sizes (B = 64, 128 x 128), where the CPU oracle would take
minutes: two independent kernels for the same op must agree, linear ops must be linear, normalised outputs must be
normalised, a training step must be deterministic.  Small-size parity against the oracle / golden vectors lives in
test_gpu_conv.py and test_gpu_model.py."""
import pytest
import torch

from helpers import canonical_model, rel_err

pytestmark = pytest.mark.gpu
B, R = 64, 128
GN_FUSE_DEFAULT = 1      # csrc/api.cu g_opt_gn_fuse


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("cin,cout,res", [(32, 32, 128), (64, 64, 64), (96, 32, 128), (256, 256, 16)])
def test_conv_two_kernels_agree_and_are_linear(cin, cout, res):
    """conv_halo_kernel (one halo box per tile, taps as operand offsets) vs conv_tc_kernel (one TMA box per tap):
    different data paths, same products -> equal up to f32 summation order; and conv(a + b) == conv(a) + conv(b)."""
    import _native as N
    g = _gen(1)
    x = torch.randn(B, res, res, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(9, cout, cin, device="cuda", generator=g) / (9 * cin) ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    out = {}
    try:
        for opt in (0, 1):
            N.lib().pub_debug_option(b"conv_halo", opt)
            out[opt] = N.conv2d_nhwc(x, w, b, ksize=3).float()
        torch.cuda.synchronize()
    finally:
        N.lib().pub_debug_option(b"conv_halo", 1)
    assert rel_err(out[1], out[0]) < 4e-3          # both round the same f32 sums to bf16
    x2 = torch.randn(B, res, res, cin, device="cuda", generator=g).bfloat16()
    lhs = N.conv2d_nhwc((x.float() + x2.float()).bfloat16(), w, None, ksize=3).float()
    rhs = N.conv2d_nhwc(x, w, None, ksize=3).float() + N.conv2d_nhwc(x2, w, None, ksize=3).float()
    assert rel_err(lhs, rhs) < 1e-2


@pytest.mark.parametrize("cin,cout,res", [(32, 32, 128), (64, 64, 64), (256, 256, 16)])
def test_wgrad_two_producers_agree_and_bias_is_column_sum(cin, cout, res):
    """box3 (one halo box, taps as row offsets) vs nine tap boxes; the bias gradient summed inside the kernel equals
    the column sums of dy; deterministic (bit-identical) across two runs."""
    import _native as N
    g = _gen(2)
    x = torch.randn(B, res, res, cin, device="cuda", generator=g).bfloat16()
    dy = torch.randn(B, res, res, cout, device="cuda", generator=g).bfloat16()
    out = {}
    try:
        for opt in (0, 1):
            N.lib().pub_debug_option(b"wgrad_box3", opt)
            out[opt] = N.conv2d_wgrad_nhwc(x, dy, 3)
        again = N.conv2d_wgrad_nhwc(x, dy, 3)
        torch.cuda.synchronize()
    finally:
        N.lib().pub_debug_option(b"wgrad_box3", 1)
    assert rel_err(out[1][0], out[0][0]) < 1e-4
    assert rel_err(out[1][1], dy.float().sum(dim=(0, 1, 2))) < 1e-4
    assert torch.equal(again[0], out[1][0]) and torch.equal(again[1], out[1][1])


def test_wgrad_is_repeatable_while_another_stream_keeps_the_sms_busy():
    """The bias sums inside the weight-gradient kernel read the staged dy tiles with ordinary shared-memory loads and
    then release the stage.  With a lane-0-only release a lane that left the mbarrier polling loop late could read a
    stage the TMA was already refilling (found with an experimental kernel variant: 3 % of its runs differed in one
    warp's 16 bias channels while a second stream ran the same kernel; profiles/r02_wgrad_rows128.txt).  Every lane
    releases now; this keeps watching the shipped kernel under the same conditions."""
    import _native as N
    g = _gen(4)
    mk = lambda: torch.randn(B, 64, 64, 64, device="cuda", generator=g).bfloat16()
    x, dy, sx, sdy = mk(), mk(), mk(), mk()
    side = torch.cuda.Stream()
    ref = N.conv2d_wgrad_nhwc(x, dy, 3)
    torch.cuda.synchronize()
    bad = 0
    for _ in range(100):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(6):
                N.conv2d_wgrad_nhwc(sx, sdy, 3)
        outs = [N.conv2d_wgrad_nhwc(x, dy, 3) for _ in range(4)]
        torch.cuda.synchronize()
        bad += sum(not (torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1])) for o in outs)
    assert bad == 0, f"{bad}/400 runs differ"


def test_full_size_training_step_is_deterministic_and_finite():
    """BASELINE configs[2] shape: same seed -> bit-identical loss and gradients (no float atomics anywhere), dropout
    keeps ~90 %, every one of the 391 gradients is finite."""
    import _native as N
    from climex_synth import make_fields
    m = canonical_model(compute_dtype="bf16", device="cuda")
    m.train()
    f = make_fields(B, R, R, 16, seed=5)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    eps = torch.randn(15, B, 32, generator=torch.Generator().manual_seed(3)).cuda()
    runs = []
    for _ in range(2):
        N.manual_seed(77)
        m.zero_grad(set_to_none=True)
        total, recon, kl = m.elbo(x, y, None, M=15, eps=eps)
        total.backward()
        runs.append((float(total), [p.grad.clone() for p in m.parameters()]))
    assert runs[0][0] == runs[1][0]
    names = [n for n, _ in m.named_parameters()]
    bad = [n for n, g0, g1 in zip(names, runs[0][1], runs[1][1]) if not torch.equal(g0, g1)]
    assert not bad, f"gradients differ between two identical steps: {bad}"
    assert all(bool(torch.isfinite(g0).all()) for g0 in runs[0][1])
    eng = m.unet.engine()
    key = [k for k in eng.block_keys if not k.endswith("_conv")][0]
    keep = eng.dropout_mask(key, B, R, R, 77).float().mean()
    assert abs(float(keep) - 0.9) < 2e-3


def test_programmatic_dependent_launch_changes_nothing_but_timing():
    """Every conv / GroupNorm / elementwise kernel is launched with programmatic stream serialisation (its prologue and,
    in the halo conv, the resident-weight TMA run before the previous kernel has finished).  A kernel reading data
    before its producer has written it would show up here: the full-size training step with the attribute on must be
    BIT-identical (loss and all 391 gradients) to the same step with ordinary stream serialisation."""
    import _native as N
    from climex_synth import make_fields
    m = canonical_model(compute_dtype="bf16", device="cuda")
    m.train()
    f = make_fields(B, R, R, 16, seed=6)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    eps = torch.randn(15, B, 32, generator=torch.Generator().manual_seed(4)).cuda()
    runs = {}
    try:
        for pdl in (0, 1, 1):
            N.lib().pub_debug_option(b"pdl", pdl)
            N.manual_seed(91)
            m.zero_grad(set_to_none=True)
            total, recon, kl = m.elbo(x, y, None, M=15, eps=eps)
            total.backward()
            torch.cuda.synchronize()
            cur = (float(total.detach()), [p.grad.clone() for p in m.parameters()])
            if pdl in runs:
                assert cur[0] == runs[pdl][0]
            runs[pdl] = cur
    finally:
        N.lib().pub_debug_option(b"pdl", 1)
    assert runs[0][0] == runs[1][0]
    for g0, g1 in zip(runs[0][1], runs[1][1]):
        assert torch.equal(g0, g1)


def test_full_size_ensemble_crps_properties():
    """100 prior members per field: CRPS >= 0, CRPS of an ensemble whose members all equal the truth is 0, and the
    CRPS kernel is invariant under a permutation of the members."""
    import metrics
    g = torch.Generator().manual_seed(8)
    T, M = 8, 100
    hr = torch.randn(T, 3, R, R, generator=g)
    ens = hr.unsqueeze(1) + 0.3 * torch.randn(T, M, 3, R, R, generator=g)
    means, arrays = metrics.crps_over_groundtruth(hr, ens)
    assert all(v >= 0 for v in means.values())
    perm = torch.randperm(M, generator=g)
    means_p, _ = metrics.crps_over_groundtruth(hr, ens[:, perm])
    for k in means:
        assert abs(means[k] - means_p[k]) < 1e-6 * max(1.0, abs(means[k]))
    zero, _ = metrics.crps_over_groundtruth(hr, hr.unsqueeze(1).expand(T, 4, 3, R, R).contiguous())
    assert all(abs(v) < 1e-7 for v in zero.values())


@pytest.mark.parametrize("name", ["bf16", "fp32"])
def test_full_resolution_training_step_matches_the_oracle_on_the_gpu(name):
    """BASELINE configs[2] at its real resolution (128 x 128, afCRPS M = 15, train-mode dropout) and B = 16: loss, CRPS,
    KL and EVERY gradient element-wise against the oracle evaluated on the SAME GPU in fp32 (TF32 off: cuDNN / cuBLAS
    IEEE paths), with the engine's dropout masks exported and injected.  The oracle is pinned on CPU by the golden
    vectors of the real reference; here it only moves to a device where 128 x 128 x 16 takes a second."""
    import _native as N
    from climex_synth import make_fields
    from oracle import probunet_oracle as O
    Bf = 16
    tol = {"fp32": 1e-4, "bf16": 1e-2}[name]
    gtol = {"fp32": 2e-3, "bf16": 1.2e-1}[name]
    cfg = O.ProbUNetCfg()
    m = canonical_model(compute_dtype=name, device="cuda")
    m.train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    f = make_fields(Bf, R, R, 16, seed=11)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    eps = torch.randn(15, Bf, 32, generator=torch.Generator().manual_seed(12)).cuda()
    N.manual_seed(4242)
    m.zero_grad(set_to_none=True)
    total, recon, kl = m.elbo(x, y, None, M=15, eps=eps)
    total.backward()
    eng = m.unet.engine()
    keys = [k for k in eng.block_keys if not k.endswith("_conv")]
    masks = {k: eng.dropout_mask(k, Bf, R, R, eng.last_seed) for k in keys}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "resample_filter" not in k}
        full = dict(sd); full.update(leaves)
        ref = O.elbo(full, cfg, x, y, eps, "afcrps", drop_masks=masks)
        ref[0].backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(float(total) - float(ref[0])) / abs(float(ref[0])) < tol, (float(total), float(ref[0]))
    assert abs(recon[0] - float(ref[1])) / abs(float(ref[1])) < 5e-3            # CRPS within 0.5 %
    assert rel_err(kl, ref[2]) < tol
    bad, worst = [], (0.0, None)
    for n, p in m.named_parameters():
        r = leaves[n].grad
        if r is None or float(r.norm()) < 1e-7:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0 or r is not None, n
            continue
        e = rel_err(p.grad, r)
        worst = max(worst, (e, n))
        if e > gtol:
            bad.append((n, e))
    assert not bad, (len(bad), worst, bad[:8])


def test_fused_groupnorm_epilogues_match_the_separate_passes():
    """GroupNorm statistics come from the epilogue of the conv that writes the tensor, and the GroupNorm-backward
    prologue (du = g * keep/(1-p) * silu'(a x + b) and its two reductions) from the epilogue of the data-gradient conv
    (pub_debug_option "gn_fuse" = 2; 1 = statistics only).  Against the separate statistics / backward passes (gn_fuse = 0) on the
    full-size training step: same loss to f32 summation-order noise, every gradient within bf16 rounding of du."""
    import _native as N
    from climex_synth import make_fields
    m = canonical_model(compute_dtype="bf16", device="cuda")
    m.train()
    f = make_fields(B, R, R, 16, seed=13)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    eps = torch.randn(15, B, 32, generator=torch.Generator().manual_seed(14)).cuda()
    runs, launches = {}, {}
    try:
        for opt in (0, 2):
            N.lib().pub_debug_option(b"gn_fuse", opt)
            N.manual_seed(123)
            m.zero_grad(set_to_none=True)
            l0 = N.lib().pub_launch_count()
            total, recon, kl = m.elbo(x, y, None, M=15, eps=eps)
            total.backward()
            torch.cuda.synchronize()
            launches[opt] = N.lib().pub_launch_count() - l0
            runs[opt] = (float(total.detach()), {n: p.grad.clone() for n, p in m.named_parameters()})
    finally:
        N.lib().pub_debug_option(b"gn_fuse", GN_FUSE_DEFAULT)
    assert launches[2] < launches[0] - 100, launches            # 57 statistics + 57 backward-reduction launches are gone
    assert abs(runs[2][0] - runs[0][0]) < 1e-4 * abs(runs[0][0]), (runs[2][0], runs[0][0])
    bad = [(n, rel_err(g, runs[0][1][n])) for n, g in runs[2][1].items()
           if float(runs[0][1][n].norm()) > 1e-7 and rel_err(g, runs[0][1][n]) > 3e-2]
    assert not bad, bad[:8]


def test_cuda_graph_training_step_matches_eager_steps():
    """graph.GraphedTrainStep: the whole step (elbo + backward + fused AdamW) captured once and replayed.
    (a) deterministic setting (eval(): dropout off, injected eps): parameters after warm-up + capture + 3 replays are
    BIT-identical to the same number (warm-up + replays) of eager steps through the same device-step optimizer path; the step count
    lives on the device.  (b) train(): the device salt gives every replay new dropout masks and eps."""
    import ctypes as C
    import _native as N
    from climex_synth import make_fields
    from graph import GraphedTrainStep
    from optim import FusedAdamW
    Bs, M = 8, 4
    f = make_fields(Bs, R, R, 16, seed=21)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    eps = torch.randn(M, Bs, 32, generator=torch.Generator().manual_seed(22)).cuda()

    def fresh(train):
        m = canonical_model(compute_dtype="bf16", device="cuda")
        m.train(train)
        return m, FusedAdamW(m.parameters(), lr=1e-3)

    m1, o1 = fresh(False)
    g = GraphedTrainStep(m1, o1, x, y, M=M, warmup=2, eps=eps)
    for _ in range(3):
        out = g(x, y)
    torch.cuda.synchronize()
    n_steps = int(g.counters[0].item())
    assert n_steps == 2 + 3                            # 2 eager warm-up steps + 3 replays (the capture pass only records)
    assert g.launches_per_step > 500
    loss_g = float(out[0])
    g.close()
    assert o1.param_groups[0]["step"] == n_steps        # close() hands the count back to the host
    m2, o2 = fresh(False)
    counters = torch.zeros(2, device="cuda", dtype=torch.int32)
    o2.use_device_step(counters[0:1])
    m2.sync_scalars = False
    for _ in range(n_steps):
        N.lib().pub_advance_counters(N.ptr(counters), N.stream())
        o2.zero_grad(set_to_none=True)
        out2 = m2.elbo(x, y, None, M=M, eps=eps)
        out2[0].backward()
        o2.step()
    torch.cuda.synchronize()
    assert int(counters[0].item()) == n_steps
    for (n1, p), q in zip(m1.named_parameters(), m2.parameters()):
        assert torch.equal(p, q), n1
    assert loss_g == float(out2[0])
    # (b) randomness under replay
    m3, o3 = fresh(True)
    g3 = GraphedTrainStep(m3, o3, x, y, M=M, warmup=1)
    a = float(g3(x, y)[0]); b = float(g3(x, y)[0])
    with torch.no_grad():
        before = [p.detach().clone() for p in m3.parameters()]
    g3.close()
    assert a != b and all(bool(torch.isfinite(p).all()) for p in before)

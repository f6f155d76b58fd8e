"""Op-level parity of the convolution family (C ABI: pub_conv2d_forward / pub_conv2d_wgrad).

Reference op: torch.nn.functional.conv2d in fp32 (TF32 off) -- the call the reference makes at
src/networks.py:89 and src/prob_unet.py:41-46 -- and its autograd weight/bias gradients.
Tolerances: fp32 path rel-err <= 1e-4, bf16 path <= 1e-2 (BASELINE.json north_star).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _setup():
    import _native as N
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return N


def _ref_conv(x_nhwc, w, b, res=None, relu=False, mask=None):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w, b, padding=w.shape[-1] // 2)
    if res is not None:
        y = y + res.float().permute(0, 3, 1, 2)
    if relu:
        y = F.relu(y)
    if mask is not None:
        y = y * (mask.float().permute(0, 3, 1, 2) > 0)
    return y.permute(0, 2, 3, 1)


CASES = [
    # B, H, W, c0, c1, cout, ks
    (2, 16, 16, 3, 0, 32, 3),
    (2, 16, 16, 6, 0, 32, 3),
    (2, 32, 32, 32, 0, 32, 3),
    (2, 32, 32, 64, 0, 64, 3),
    (3, 16, 16, 128, 0, 128, 3),
    (2, 16, 16, 256, 0, 256, 3),
    (2, 16, 16, 256, 256, 256, 3),
    (2, 16, 16, 256, 128, 256, 1),
    (2, 32, 32, 64, 32, 32, 3),
    (2, 32, 32, 32, 0, 64, 1),
    (3, 8, 8, 256, 0, 256, 3),      # 128-pixel patch spans two images, odd batch -> OOB batch rows
    (2, 16, 16, 256, 0, 512, 3),    # N tiled 2 x 256 (data-gradient of a 512-channel concat input)
    (2, 16, 16, 128, 0, 384, 3),    # N tiled 3 x 128
    (1, 128, 128, 32, 0, 32, 3),
    (5, 4, 4, 64, 0, 64, 3),
    (2, 32, 16, 64, 0, 64, 3),      # non-square: 2 x 2 halo tiles per image
    (1, 16, 48, 32, 32, 64, 3),     # non-square + virtual concat (two activation tensor maps)
    (2, 48, 24, 96, 0, 32, 3),      # 96 channels -> 64-byte-row K blocks, 3 x 3 tiles
]


def to_tf32(x):
    """round-to-nearest onto the tf32 grid (10-bit mantissa), what PUB_TF32 tensors hold"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


# the A/B tests force the per-tap kernels, whose 128-pixel patches do not tile 48 x 24 (the halo / box3 kernels do;
# that shape is checked against the fp32 reference by the plain forward / wgrad tests)
_TAP_TILEABLE = lambda c: (c[1], c[2]) != (48, 24)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_tf32"])
def test_conv_forward(case, mode):
    N = _setup()
    B, H, W, c0, c1, cout, ks = case
    dt = torch.bfloat16 if mode.endswith("bf16") else torch.float32
    backend = N.BACKEND_TCGEN05 if mode.startswith("tc") else N.BACKEND_SIMT
    ndt = {"simt_f32": N.F32, "simt_bf16": N.BF16, "tc_bf16": N.BF16, "tc_tf32": N.TF32}[mode]
    cast = to_tf32 if mode == "tc_tf32" else (lambda t: t.to(dt))
    if mode.startswith("tc") and (c0 % 32 or c1 % 32):
        pytest.skip("tcgen05 path needs Cin % 32 == 0 (first layers run on the SIMT kernel)")
    g = torch.Generator(device="cuda").manual_seed(1)
    # inputs are channel slices of wider buffers -> exercises the pixel-stride (ld) handling
    buf0 = cast(torch.randn(B, H, W, c0 + 32, device="cuda", generator=g))
    x0 = buf0[..., :c0]
    x1 = cast(torch.randn(B, H, W, c1, device="cuda", generator=g)) if c1 else None
    w = torch.randn(cout, c0 + c1, ks, ks, device="cuda", generator=g) / ((c0 + c1) * ks * ks) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    res = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    mask = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    wp = N.pack_conv_weight(w, ndt)
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=3)
    wq = wp.float().permute(1, 2, 0).reshape(cout, c0 + c1, ks, ks)   # the (rounded) weights the kernel sees
    tol = {"simt_f32": 1e-4, "simt_bf16": 1e-2, "tc_bf16": 1e-2, "tc_tf32": 1e-3}[mode]   # tf32: output rounded to 10 bits
    for kw in (dict(), dict(res=res), dict(relu=True), dict(res=res, mask=mask)):
        y = N.conv2d_nhwc(x0, wp, b, x1=x1, ksize=ks, backend=backend, dtype=ndt, **kw)
        ref = _ref_conv(xin, wq, b, **kw)
        torch.cuda.synchronize()
        e = rel_err(y.float(), ref)
        assert e < tol, f"{mode} {case} {list(kw)} rel_err={e:.3e}"


@pytest.mark.parametrize("case", [c for c in CASES if c[0] * c[1] * c[2] <= 4096 * 4])
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_tf32"])
def test_conv_wgrad(case, mode):
    N = _setup()
    B, H, W, c0, c1, cout, ks = case
    dt = torch.bfloat16 if mode.endswith("bf16") else torch.float32
    backend = N.BACKEND_TCGEN05 if mode.startswith("tc") else N.BACKEND_SIMT
    ndt = {"simt_f32": N.F32, "simt_bf16": N.BF16, "tc_bf16": N.BF16, "tc_tf32": N.TF32}[mode]
    cast = to_tf32 if mode == "tc_tf32" else (lambda t: t.to(dt))
    if mode.startswith("tc") and (c0 % 32 or c1 % 32 or (H * W) % 64):
        pytest.skip("tcgen05 wgrad needs Cin % 32 == 0 and 64-pixel K tiles")
    g = torch.Generator(device="cuda").manual_seed(2)
    x0 = cast(torch.randn(B, H, W, c0, device="cuda", generator=g))
    x1 = cast(torch.randn(B, H, W, c1, device="cuda", generator=g)) if c1 else None
    dy = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    dw, db = N.conv2d_wgrad_nhwc(x0, dy, ks, x1=x1, backend=backend, dtype=ndt)
    xin = (x0 if x1 is None else torch.cat([x0, x1], dim=3)).float().permute(0, 3, 1, 2)
    w = torch.zeros(cout, c0 + c1, ks, ks, device="cuda", requires_grad=True)
    bb = torch.zeros(cout, device="cuda", requires_grad=True)
    F.conv2d(xin, w, bb, padding=ks // 2).backward(dy.float().permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    tol = 1e-4 if dt == torch.float32 else 2e-3   # inputs are identical bf16 values; only f32 summation order differs
    assert rel_err(dw, w.grad) < tol, f"{mode} {case} dw rel_err={rel_err(dw, w.grad):.3e}"
    assert rel_err(db, bb.grad) < tol, f"{mode} {case} db rel_err={rel_err(db, bb.grad):.3e}"


def test_dgrad_via_transposed_pack():
    """Data gradient = forward kernel on mirrored/transposed weights (pub_pack_conv_weight transpose_flip=1)."""
    N = _setup()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, cin, cout = 2, 16, 16, 64, 128
    x = torch.randn(B, cin, H, W, device="cuda", generator=g, requires_grad=True)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / 24
    dy = torch.randn(B, H, W, cout, device="cuda", generator=g)
    F.conv2d(x, w, padding=1).backward(dy.permute(0, 3, 1, 2))
    for dt, backend, tol in ((torch.float32, N.BACKEND_SIMT, 1e-4), (torch.bfloat16, N.BACKEND_SIMT, 1e-2),
                             (torch.bfloat16, N.BACKEND_TCGEN05, 1e-2)):
        wt = N.pack_conv_weight(w, N.BF16 if dt == torch.bfloat16 else N.F32, transpose_flip=True)
        dx = N.conv2d_nhwc(dy.to(dt), wt, None, ksize=3, backend=backend)
        torch.cuda.synchronize()
        e = rel_err(dx.float().permute(0, 3, 1, 2), x.grad)
        assert e < tol, f"dgrad {dt} backend={backend} rel_err={e:.3e}"


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin", [3, 6])
def test_first_layer_small_cin_kernels(dt, cin):
    """The engines stage the 3 / 6-variable network input in an 8-channel padded NHWC buffer; forward and
    weight gradient then run on the dedicated small-Cin kernels (conv_smallc_kernel / wgrad_smallc_kernel)."""
    N = _setup()
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, W, cout = 3, 32, 32, 32
    buf = torch.randn(B, H, W, 8, device="cuda", generator=g).to(dt)
    buf[..., cin:] = float("nan")                      # padding channels must never be read into the result
    x0 = buf[..., :cin]
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    ndt = N.BF16 if dt == torch.bfloat16 else N.F32
    wp = N.pack_conv_weight(w, ndt)
    wq = wp.float().permute(1, 2, 0).reshape(cout, cin, 3, 3)
    y = N.conv2d_nhwc(x0, wp, b, ksize=3, relu=True, backend=N.BACKEND_SIMT)
    ref = _ref_conv(x0, wq, b, relu=True)
    tol = 1e-4 if dt == torch.float32 else 1e-2
    assert rel_err(y.float(), ref) < tol
    dy = torch.randn(B, H, W, cout, device="cuda", generator=g).to(dt)
    dw, db = N.conv2d_wgrad_nhwc(x0, dy, 3, backend=N.BACKEND_SIMT)
    wz = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
    bz = torch.zeros(cout, device="cuda", requires_grad=True)
    F.conv2d(x0.float().permute(0, 3, 1, 2), wz, bz, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    assert rel_err(dw, wz.grad) < (1e-4 if dt == torch.float32 else 2e-3)
    assert rel_err(db, bz.grad) < (1e-4 if dt == torch.float32 else 2e-3)


@pytest.mark.parametrize("mode", ["tc_bf16", "tc_tf32"])
@pytest.mark.parametrize("case", [c for c in CASES if c[6] == 3 and c[1] % 16 == 0 and c[2] % 8 == 0 and c[3] % 32 == 0
                                  and _TAP_TILEABLE(c)])
def test_conv_halo_kernel_on_every_legal_shape(case, mode):
    """By default only the shapes where it measured faster go to conv_halo_kernel (one halo load per tile, taps as
    start-address offsets of the swizzled operand); force it on every legal 3x3 shape, residual + mask epilogue."""
    N = _setup()
    B, H, W, c0, c1, cout, ks = case
    ndt = N.BF16 if mode == "tc_bf16" else N.TF32
    cast = to_tf32 if mode == "tc_tf32" else (lambda t: t.to(torch.bfloat16))
    g = torch.Generator(device="cuda").manual_seed(11)
    x0 = cast(torch.randn(B, H, W, c0 + 32, device="cuda", generator=g))[..., :c0]
    x1 = cast(torch.randn(B, H, W, c1, device="cuda", generator=g)) if c1 else None
    w = torch.randn(cout, c0 + c1, ks, ks, device="cuda", generator=g) / ((c0 + c1) * ks * ks) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    res = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    mask = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    wp = N.pack_conv_weight(w, ndt)
    wq = wp.float().permute(1, 2, 0).reshape(cout, c0 + c1, ks, ks)
    xin = x0 if x1 is None else torch.cat([x0, x1], dim=3)
    ref = _ref_conv(xin, wq, b, res=res, mask=mask)
    out = {}
    try:
        for opt in (0, 2):
            N.lib().pub_debug_option(b"conv_halo", opt)
            out[opt] = N.conv2d_nhwc(x0, wp, b, x1=x1, ksize=ks, backend=N.BACKEND_TCGEN05, dtype=ndt, res=res, mask=mask)
            torch.cuda.synchronize()
    finally:
        N.lib().pub_debug_option(b"conv_halo", 1)
    tol = 1e-2 if mode == "tc_bf16" else 1e-3
    assert rel_err(out[2].float(), ref) < tol and rel_err(out[0].float(), ref) < tol
    assert rel_err(out[2].float(), out[0].float()) < 1e-2 * tol + (4e-3 if mode == "tc_bf16" else 2e-4)   # same products, other order


@pytest.mark.parametrize("mode", ["tc_bf16", "tc_tf32"])
@pytest.mark.parametrize("case", [c for c in CASES if c[6] == 3 and c[1] % 8 == 0 and c[2] % 8 == 0 and c[3] % 32 == 0
                                  and _TAP_TILEABLE(c)
                                  and c[0] * c[1] * c[2] <= 4096 * 4])
def test_wgrad_box3_and_tap_boxes_agree(case, mode):
    """3x3 weight gradient: three (8+2) x 8 boxes with the horizontal taps as row offsets of the swizzled operand
    (default) against the nine-tap-box producer and against autograd."""
    N = _setup()
    B, H, W, c0, c1, cout, ks = case
    ndt = N.BF16 if mode == "tc_bf16" else N.TF32
    cast = to_tf32 if mode == "tc_tf32" else (lambda t: t.to(torch.bfloat16))
    g = torch.Generator(device="cuda").manual_seed(12)
    x0 = cast(torch.randn(B, H, W, c0, device="cuda", generator=g))
    x1 = cast(torch.randn(B, H, W, c1, device="cuda", generator=g)) if c1 else None
    dy = cast(torch.randn(B, H, W, cout, device="cuda", generator=g))
    out = {}
    try:
        for opt in (0, 1):
            N.lib().pub_debug_option(b"wgrad_box3", opt)
            out[opt] = N.conv2d_wgrad_nhwc(x0, dy, ks, x1=x1, backend=N.BACKEND_TCGEN05, dtype=ndt)
            torch.cuda.synchronize()
    finally:
        N.lib().pub_debug_option(b"wgrad_box3", 1)
    xin = (x0 if x1 is None else torch.cat([x0, x1], dim=3)).float().permute(0, 3, 1, 2)
    w = torch.zeros(cout, c0 + c1, ks, ks, device="cuda", requires_grad=True)
    F.conv2d(xin, w, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    tol = 2e-3
    assert rel_err(out[1][0], w.grad) < tol, f"box3 {mode} {case}: {rel_err(out[1][0], w.grad):.3e}"
    assert rel_err(out[0][0], w.grad) < tol, f"tap boxes {mode} {case}: {rel_err(out[0][0], w.grad):.3e}"
    assert rel_err(out[1][1], out[0][1]) < 1e-6

"""ctypes binding of libprobunet_b200.so (the C ABI in include/probunet_b200.h) + the autograd glue.

PyTorch is used here for device memory, streams and autograd bookkeeping only: every
tensor handed to the library is a torch CUDA allocation passed by ``data_ptr()`` together
with ``torch.cuda.current_stream()``; all arithmetic happens in the library's sm_100a
kernels.  There is deliberately NO fallback: if the shared library is missing or a tensor
is not on a CUDA device the call raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libprobunet_b200.so")
_lib = None

F32, BF16, TF32, TF32_BF16S0 = 0, 1, 2, 3
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2
_DTYPE_NAMES = {"fp32": F32, "float32": F32, "f32": F32, "bf16": BF16, "bfloat16": BF16, "tf32": TF32,
                "tf32_bf16s0": TF32_BF16S0}
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, TF32: torch.float32}

# process-wide knobs (env vars so that the reference's drivers stay unchanged)
_default_dtype = os.environ.get("PROBUNET_B200_DTYPE", "bf16")
_backend = {"auto": 0, "simt": 1, "tcgen05": 2}[os.environ.get("PROBUNET_B200_BACKEND", "auto")]


class NativeError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("c0", C.c_int32), ("ld0", C.c_int32),
                ("x1", C.c_void_p), ("c1", C.c_int32), ("ld1", C.c_int32),
                ("w", C.c_void_p), ("bias", C.c_void_p),
                ("res", C.c_void_p), ("ld_res", C.c_int32),
                ("mask", C.c_void_p), ("ld_mask", C.c_int32),
                ("y", C.c_void_p), ("ldy", C.c_int32),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("relu", C.c_int32), ("dtype", C.c_int32), ("backend", C.c_int32)]


class ConvGnArgs(C.Structure):
    _fields_ = [("stat_part", C.c_void_p), ("gn_bwd", C.c_int32),
                ("gx0", C.c_void_p), ("gx1", C.c_void_p), ("gc0", C.c_int32), ("gld0", C.c_int32), ("gld1", C.c_int32),
                ("gcoef", C.c_void_p), ("p_drop", C.c_float), ("seed", C.c_uint64), ("subseq", C.c_uint64)]


class WgradArgs(C.Structure):
    _fields_ = [("x0", C.c_void_p), ("c0", C.c_int32), ("ld0", C.c_int32),
                ("x1", C.c_void_p), ("c1", C.c_int32), ("ld1", C.c_int32),
                ("dy", C.c_void_p), ("ld_dy", C.c_int32),
                ("dw", C.c_void_p), ("dbias", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("accumulate", C.c_int32), ("dtype", C.c_int32), ("backend", C.c_int32)]


class UNetBlockDesc(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("up", C.c_int32), ("down", C.c_int32),
                ("has_skip_conv", C.c_int32), ("is_conv", C.c_int32)]


class FcombArgs(C.Structure):
    _fields_ = [("feat", C.c_void_p), ("feat_nchw", C.c_int32), ("dtype", C.c_int32),
                ("stride", C.c_int64 * 4), ("z", C.c_void_p),
                ("w0", C.c_void_p), ("b0", C.c_void_p), ("w1", C.c_void_p), ("b1", C.c_void_p),
                ("w2", C.c_void_p), ("b2", C.c_void_p), ("out", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("F", C.c_int32), ("L", C.c_int32),
                ("C", C.c_int32), ("M", C.c_int32)]


class AdamWEntry(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


# every symbol include/probunet_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "pub_last_error", "pub_version", "pub_launch_count", "pub_debug_option", "pub_debug_pointer", "pub_conv2d_forward", "pub_pack_conv_weight", "pub_conv2d_wgrad_workspace",
    "pub_conv2d_wgrad", "pub_nchw_to_nhwc", "pub_nhwc_to_nchw", "pub_unet_create", "pub_unet_destroy",
    "pub_unet_num_params", "pub_unet_workspace_bytes", "pub_unet_forward", "pub_unet_backward",
    "pub_unet_dropout_mask", "pub_encoder_create", "pub_encoder_destroy", "pub_encoder_num_params",
    "pub_encoder_workspace_bytes", "pub_encoder_forward", "pub_encoder_backward", "pub_rsample_forward",
    "pub_rsample_backward", "pub_kl_normal_forward", "pub_kl_normal_backward", "pub_fcomb_forward_workspace", "pub_fcomb_forward",
    "pub_fcomb_backward_workspace", "pub_fcomb_backward", "pub_loss_workspace", "pub_ensemble_loss", "pub_l1_loss", "pub_msssim_workspace", "pub_wmse_msssim_loss", "pub_climex_stats", "pub_climex_transform",
    "pub_scale_by_device_scalar", "pub_ensemble_metrics_workspace", "pub_ensemble_metrics", "pub_adamw_step",
    "pub_groupnorm_scratch_bytes", "pub_groupnorm_silu_forward", "pub_groupnorm_silu_backward",
    "pub_conv2d_fused_rows", "pub_conv2d_forward_fused", "pub_advance_counters", "pub_adamw_step_dev",
    "pub_psd_table_ints", "pub_psd_build_table", "pub_radial_psd", "pub_histogram",
]


def lib():
    """Loads the shared library (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU / PyTorch fallback for the Prob U-Net hot path)")
        l = C.CDLL(_LIB_PATH)
        l.pub_last_error.restype = C.c_char_p
        l.pub_launch_count.restype = C.c_ulonglong
        for name in ("pub_conv2d_wgrad_workspace", "pub_unet_workspace_bytes", "pub_encoder_workspace_bytes",
                     "pub_fcomb_backward_workspace", "pub_loss_workspace", "pub_fcomb_forward_workspace",
                     "pub_ensemble_metrics_workspace", "pub_msssim_workspace", "pub_groupnorm_scratch_bytes", "pub_psd_table_ints"):
            if hasattr(l, name):
                getattr(l, name).restype = C.c_size_t
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise NativeError(f"{what} failed ({rc}): {lib().pub_last_error().decode()}")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def resolve_dtype(name):
    return _DTYPE_NAMES[(name or _default_dtype).lower()]


def resolve_encoder_dtype(name):
    """Precision of the two Gaussian encoders.  Their log-sigma head feeds exp(): an absolute error of d in
    log sigma is a RELATIVE error d in sigma and in z = mu + sigma*eps, so with the U-Net in bf16 the encoders run
    one notch higher (tf32 tensor cores, f32 storage).  PROBUNET_B200_ENCODER_DTYPE (fp32 | tf32 | tf32_bf16s0 |
    bf16) overrides it; tf32_bf16s0 keeps only the full-resolution first stage in bf16 (forward error 2.5x tf32's, 1 %
    faster, noisier first-stage gradients: DESIGN.md, precision policy)."""
    dt = resolve_dtype(name)
    env = os.environ.get("PROBUNET_B200_ENCODER_DTYPE")
    if env:
        return _DTYPE_NAMES[env.lower()]
    return TF32 if dt == BF16 else dt


def set_default_dtype(name):
    global _default_dtype
    assert name.lower() in _DTYPE_NAMES
    _default_dtype = name


def set_backend(name):
    global _backend
    _backend = {"auto": 0, "simt": 1, "tcgen05": 2}[name]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NativeError("probunet_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")


def _ptr_table(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


# ======================================================================================
# primitive ops (used by the op-level tests and by the standalone networks.Conv2d.forward)
# ======================================================================================
def pack_conv_weight(w, dtype, transpose_flip=False):
    cout, cin, k, _ = w.shape
    n, kk = (cin, cout) if transpose_flip else (cout, cin)
    out = torch.empty(k * k, n, kk, device=w.device, dtype=_TORCH_DT[dtype])
    check(lib().pub_pack_conv_weight(ptr(w.contiguous()), ptr(out), cout, cin, k, dtype, int(transpose_flip), stream()),
          "pub_pack_conv_weight")
    return out


def conv2d_nhwc(x0, w_packed, bias=None, x1=None, res=None, mask=None, relu=False, ksize=3, backend=None, out=None,
                dtype=None, gn_stats=False, gn_bwd=None):
    """x0/x1/res/mask: NHWC views [B,H,W,C] (last-dim-contiguous, arbitrary pixel stride).

    gn_stats=True: also returns the GroupNorm statistics partials [B, rows, cout, 2] the epilogue emitted.
    gn_bwd=dict(x0=, x1=None, coef=, p_drop=0.0, seed=0, subseq=0): a data-gradient launch with the GroupNorm-backward
    prologue fused (returns (du, partials)); see pub_conv2d_forward_fused."""
    require_cuda(x0, w_packed)
    B, H, W, c0 = x0.shape
    dt = dtype if dtype is not None else (BF16 if x0.dtype == torch.bfloat16 else F32)
    cout = w_packed.shape[1]
    y = out if out is not None else torch.empty(B, H, W, cout, device=x0.device, dtype=x0.dtype)
    a = ConvArgs()
    a.x0, a.c0, a.ld0 = x0.data_ptr(), c0, x0.stride(2)
    if x1 is not None:
        a.x1, a.c1, a.ld1 = x1.data_ptr(), x1.shape[3], x1.stride(2)
    a.w = w_packed.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    if res is not None:
        a.res, a.ld_res = res.data_ptr(), res.stride(2)
    if mask is not None:
        a.mask, a.ld_mask = mask.data_ptr(), mask.stride(2)
    a.y, a.ldy = y.data_ptr(), y.stride(2)
    a.B, a.H, a.W, a.cout, a.ksize = B, H, W, cout, ksize
    a.relu, a.dtype, a.backend = int(relu), dt, _backend if backend is None else backend
    if gn_stats or gn_bwd is not None:
        rows = lib().pub_conv2d_fused_rows(C.byref(a))
        if rows <= 0:
            raise NativeError("this conv launch has no fused GroupNorm epilogue (pub_conv2d_fused_rows == 0)")
        part = torch.empty(B, rows, cout, 2, device=x0.device, dtype=torch.float32)
        g = ConvGnArgs()
        g.stat_part = part.data_ptr()
        if gn_bwd is not None:
            gx0, gx1 = gn_bwd["x0"], gn_bwd.get("x1")
            g.gn_bwd, g.gx0, g.gc0, g.gld0 = 1, gx0.data_ptr(), gx0.shape[3], gx0.stride(2)
            if gx1 is not None:
                g.gx1, g.gld1 = gx1.data_ptr(), gx1.stride(2)
            g.gcoef = gn_bwd["coef"].data_ptr()
            g.p_drop, g.seed, g.subseq = float(gn_bwd.get("p_drop", 0.0)), int(gn_bwd.get("seed", 0)), int(gn_bwd.get("subseq", 0))
        check(lib().pub_conv2d_forward_fused(C.byref(a), C.byref(g), stream()), "pub_conv2d_forward_fused")
        return y, part
    check(lib().pub_conv2d_forward(C.byref(a), stream()), "pub_conv2d_forward")
    return y


def conv2d_wgrad_nhwc(x0, dy, ksize, x1=None, want_bias=True, backend=None, dtype=None):
    B, H, W, c0 = x0.shape
    dt = dtype if dtype is not None else (BF16 if x0.dtype == torch.bfloat16 else F32)
    cout = dy.shape[3]
    cin = c0 + (x1.shape[3] if x1 is not None else 0)
    dw = torch.empty(cout, cin, ksize, ksize, device=x0.device, dtype=torch.float32)
    db = torch.empty(cout, device=x0.device, dtype=torch.float32) if want_bias else None
    a = WgradArgs()
    a.x0, a.c0, a.ld0 = x0.data_ptr(), c0, x0.stride(2)
    if x1 is not None:
        a.x1, a.c1, a.ld1 = x1.data_ptr(), x1.shape[3], x1.stride(2)
    a.dy, a.ld_dy = dy.data_ptr(), dy.stride(2)
    a.dw, a.dbias = dw.data_ptr(), (db.data_ptr() if db is not None else None)
    a.B, a.H, a.W, a.cout, a.ksize = B, H, W, cout, ksize
    a.accumulate, a.dtype, a.backend = 0, dt, _backend if backend is None else backend
    nbytes = lib().pub_conv2d_wgrad_workspace(C.byref(a))
    ws = torch.empty(max(nbytes, 16), device=x0.device, dtype=torch.uint8)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    check(lib().pub_conv2d_wgrad(C.byref(a), stream()), "pub_conv2d_wgrad")
    return dw, db


def nchw_to_nhwc(x, dtype, x1=None):
    require_cuda(x)
    B, c0, H, W = x.shape
    c1 = x1.shape[1] if x1 is not None else 0
    y = torch.empty(B, H, W, c0 + c1, device=x.device, dtype=_TORCH_DT[dtype])
    check(lib().pub_nchw_to_nhwc(ptr(x.contiguous()), c0, ptr(x1.contiguous() if x1 is not None else None), c1,
                                 ptr(y), c0 + c1, B, H, W, dtype, stream()), "pub_nchw_to_nhwc")
    return y


def nhwc_to_nchw(x):
    B, H, W, Cc = x.shape
    dt = BF16 if x.dtype == torch.bfloat16 else F32
    y = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
    check(lib().pub_nhwc_to_nchw(ptr(x), x.stride(2), Cc, ptr(y), B, H, W, dt, 0, stream()), "pub_nhwc_to_nchw")
    return y


class _GroupNormSiluFn(torch.autograd.Function):
    """y = resample(dropout(silu(film(group_norm(x))))) on NHWC tensors (pub_groupnorm_silu_forward / _backward)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, film, resample, p_drop, seed, subseq):
        B, H, W, Cc = x.shape
        dt = BF16 if x.dtype == torch.bfloat16 else F32
        Ho, Wo = (H // 2, W // 2) if resample == 1 else ((H * 2, W * 2) if resample == 2 else (H, W))
        y = torch.empty(B, Ho, Wo, Cc, device=x.device, dtype=x.dtype)
        G = min(32, Cc // 4)
        stats = torch.empty(B, G, 2, device=x.device, dtype=torch.float32)
        coef = torch.empty(B, Cc, 2, device=x.device, dtype=torch.float32)
        nb = lib().pub_groupnorm_scratch_bytes(B, Cc, H, W)
        scratch = torch.empty(nb, device=x.device, dtype=torch.uint8)
        check(lib().pub_groupnorm_silu_forward(ptr(x), Cc, x.stride(2), B, H, W, ptr(gamma), ptr(beta), ptr(film), resample,
                                               C.c_float(p_drop), C.c_uint64(seed), C.c_uint64(subseq), ptr(y), ptr(stats),
                                               ptr(coef), ptr(scratch), C.c_size_t(nb), dt, stream()),
              "pub_groupnorm_silu_forward")
        ctx.save_for_backward(x, gamma, beta, film if film is not None else gamma.new_empty(0), stats, coef)
        ctx.cfg = (resample, p_drop, seed, subseq, dt, film is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, film, stats, coef = ctx.saved_tensors
        resample, p_drop, seed, subseq, dt, has_film = ctx.cfg
        B, H, W, Cc = x.shape
        dx = torch.empty(B, H, W, Cc, device=x.device, dtype=x.dtype)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
        dfilm = torch.empty(2 * Cc, device=x.device, dtype=torch.float32) if has_film else None
        nb = lib().pub_groupnorm_scratch_bytes(B, Cc, H, W)
        scratch = torch.empty(nb, device=x.device, dtype=torch.uint8)
        check(lib().pub_groupnorm_silu_backward(ptr(x), Cc, x.stride(2), B, H, W, ptr(gamma), ptr(beta),
                                                ptr(film if has_film else None), resample, C.c_float(p_drop),
                                                C.c_uint64(seed), C.c_uint64(subseq), ptr(stats), ptr(coef),
                                                ptr(dy.contiguous()), ptr(dx), ptr(dgamma), ptr(dbeta), ptr(dfilm),
                                                ptr(scratch), C.c_size_t(nb), dt, stream()), "pub_groupnorm_silu_backward")
        return dx, dgamma, dbeta, dfilm, None, None, None, None


def groupnorm_silu_nhwc(x, gamma, beta, film=None, resample=0, p_drop=0.0, seed=0, subseq=0):
    """x [B,H,W,C] NHWC (f32 or bf16, last dim contiguous) -> silu(GroupNorm(x) [* (1+scale) + shift]) [dropout] [2x]."""
    require_cuda(x, gamma, beta, film)
    if x.stride(3) != 1 or x.stride(1) != x.shape[2] * x.stride(2) or x.stride(0) != x.shape[1] * x.stride(1):
        raise ValueError("groupnorm_silu_nhwc expects an NHWC view with a uniform pixel stride")
    return _GroupNormSiluFn.apply(x, gamma.contiguous().float(), beta.contiguous().float(),
                                  film.contiguous().float() if film is not None else None, int(resample), float(p_drop),
                                  int(seed), int(subseq))


def conv2d(x, weight, bias):
    """Standalone NCHW f32 conv (networks.Conv2d.forward without resampling)."""
    dt = resolve_dtype(None)
    xh = nchw_to_nhwc(x, dt)
    y = conv2d_nhwc(xh, pack_conv_weight(weight, dt), bias, ksize=weight.shape[-1])
    return nhwc_to_nchw(y)


# ======================================================================================
# Philox stream shared by rsample / dropout (seeded from torch's global seed on first use)
# ======================================================================================
class _Rng:
    seed = None
    offset = 0

    @classmethod
    def next(cls, n=1):
        if cls.seed is None:
            # default: torch's global seed, with the data-parallel rank folded in -- every rank usually calls the same
            # torch.manual_seed (needed for identical initial weights), and identical dropout masks / eps on all
            # ranks would correlate the shards of the global batch
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()
            except Exception:
                rank = 0
            cls.seed = (torch.initial_seed() + 0x9E3779B97F4A7C15 * rank) & 0x7FFFFFFFFFFFFFFF
        o = cls.offset
        cls.offset += n
        return cls.seed, o


def manual_seed(seed):
    """Re-seeds the engine's Philox streams (dropout masks, rsample eps)."""
    _Rng.seed, _Rng.offset = int(seed) & 0x7FFFFFFFFFFFFFFF, 0


class _WorkspacePool:
    """Engine workspaces are multi-GB; allocating a fresh one per forward makes the caching allocator split and
    re-malloc huge blocks (measured: sporadic ~1 s stalls).  Buffers are checked out for forward..backward and
    returned afterwards, so a steady-state training step allocates nothing."""

    MAX_FREE = 6          # buffers kept per device; beyond that the smallest ones are released to the allocator

    def __init__(self):
        self.free = {}    # (device, stream) -> list of free buffers, most recently returned last

    @staticmethod
    def _key(device):
        # Buffers are handed back when the work that uses them has been ENQUEUED, not finished: reuse is only ordered
        # on the same stream, so every stream has its own free list (the Gaussian encoders run on side streams, an
        # ensemble pipeline may alternate streams per field batch).
        device = torch.device(device)
        if device.type != "cuda":
            return device
        return (device, torch.cuda.current_stream(device).cuda_stream)

    def take(self, nbytes, device):
        """Best fit: the smallest free buffer that is large enough (a ragged last batch or a validation batch reuses
        the training step's workspace instead of pinning another multi-GB block)."""
        lst = self.free.get(self._key(device), [])
        best = None
        for i, b in enumerate(lst):
            if b.numel() >= nbytes and (best is None or b.numel() < lst[best].numel()):
                best = i
        if best is not None:
            return lst.pop(best)
        return torch.empty(nbytes, device=device, dtype=torch.uint8)

    def give(self, ws):
        if ws is None:
            return
        lst = self.free.setdefault(self._key(ws.device), [])
        lst.append(ws)
        while len(lst) > self.MAX_FREE:
            lst.pop(min(range(len(lst)), key=lambda i: lst[i].numel()))

    def clear(self):
        """Releases every cached workspace (call when the batch shape changes for good, e.g. after training)."""
        self.free.clear()


workspaces = _WorkspacePool()


# gradient-ready callback (set by parallel.GradSynchronizer): called with each engine's flat
# gradient buffer right after its backward kernels have been enqueued
grad_ready_callback = None


def _flat_grads(params):
    """One flat f32 buffer + per-parameter views (keeps a sub-network's gradients contiguous so
    that they can be all-reduced with a single collective)."""
    total = sum(p.numel() for p in params)
    flat = torch.empty(total, device=params[0].device, dtype=torch.float32)
    views, off = [], 0
    for p in params:
        views.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    return flat, views


def _notify(flat, params=None, views=None):
    if grad_ready_callback is not None:
        grad_ready_callback(flat, params, views)


def _consume_ws(ctx, what):
    """The engines' backward passes consume the workspace saved by forward: a second backward through the same graph
    (retain_graph=True, two losses sharing the features) would hand the kernels a recycled buffer."""
    if ctx.ws is None:
        raise NativeError(f"{what}: backward was already run for this forward (retain_graph / a second backward "
                          "through the same engine call is not supported: the saved workspace is recycled)")
    ws, ctx.ws = ctx.ws, None
    return ws


# ======================================================================================
# U-Net engine
# ======================================================================================
class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, training, nhwc_out, seed, nzero, *params):
        B, _, H, W = x.shape
        L = lib()
        nbytes = L.pub_unet_workspace_bytes(eng.handle, B, H, W)
        if nbytes == 0:
            raise NativeError("pub_unet_workspace_bytes: " + L.pub_last_error().decode())
        ws = workspaces.take(nbytes, x.device)
        real = params[:len(params) - nzero]
        if nhwc_out:
            out = torch.empty(B, H, W, eng.out_ch, device=x.device, dtype=_TORCH_DT[eng.dtype])
        else:
            out = torch.empty(B, eng.out_ch, H, W, device=x.device, dtype=torch.float32)
        xc = x.contiguous()
        check(L.pub_unet_forward(eng.handle, B, H, W, ptr(xc), _ptr_table(real), ptr(out), int(not nhwc_out), ptr(ws),
                                 C.c_size_t(nbytes), C.c_uint64(seed), int(training), _backend, stream()), "pub_unet_forward")
        ctx.eng, ctx.nbytes, ctx.shape = eng, nbytes, (B, H, W)
        ctx.training, ctx.nhwc_out, ctx.seed, ctx.nzero = training, nhwc_out, seed, nzero
        ctx.x_req = ctx.needs_input_grad[1]
        if any(ctx.needs_input_grad):
            ctx.ws = ws                                  # held until backward
            ctx.save_for_backward(*params)
        else:
            ctx.ws = None
            workspaces.give(ws)                          # inference: nothing to keep
        return out

    @staticmethod
    def backward(ctx, dout):
        eng, (B, H, W) = ctx.eng, ctx.shape
        params = ctx.saved_tensors
        real = params[:len(params) - ctx.nzero]
        zero = params[len(params) - ctx.nzero:]
        flat, gviews = _flat_grads(list(real) + list(zero))
        if ctx.nzero:
            nz = sum(p.numel() for p in zero)
            flat[flat.numel() - nz:].zero_()       # map_label / affine weights: exactly-zero gradients
        dx = torch.empty(B, eng.in_ch, H, W, device=dout.device, dtype=torch.float32) if ctx.x_req else None
        dout = dout.contiguous()
        ws = _consume_ws(ctx, "UNet")
        check(lib().pub_unet_backward(eng.handle, B, H, W, ptr(dout), int(not ctx.nhwc_out), _ptr_table(real),
                                      _ptr_table(gviews[:len(real)]), ptr(dx), ptr(ws), C.c_size_t(ctx.nbytes),
                                      C.c_uint64(ctx.seed), int(ctx.training), _backend, stream()), "pub_unet_backward")
        workspaces.give(ws)
        _notify(flat, list(real) + list(zero), gviews)
        return (None, dx, None, None, None, None) + tuple(gviews)


class UNetEngine:
    """Owns the native plan of one networks.UNet module (block list + parameter order)."""

    def __init__(self, module, dtype):
        self.dtype = dtype
        self.module = module
        self.in_ch, self.out_ch = module.in_channels, module.out_channels
        enc, dec = [], []
        self.params, self.zero_params = [], []
        self.block_keys = []
        from networks import UNetBlock
        self.block_modules = []
        for name, md, dst in ([("enc." + k, v, enc) for k, v in module.enc.items()] +
                              [("dec." + k, v, dec) for k, v in module.dec.items()]):
            d = UNetBlockDesc()
            if isinstance(md, UNetBlock):
                has_conv = md.skip is not None and md.skip.weight is not None
                d.cin, d.cout, d.up, d.down = md.in_channels, md.out_channels, int(md.up), int(md.down)
                d.has_skip_conv, d.is_conv = int(has_conv), 0
                self.params += [md.norm0.weight, md.norm0.bias, md.conv0.weight, md.conv0.bias, md.affine.bias,
                                md.norm1.weight, md.norm1.bias, md.conv1.weight, md.conv1.bias]
                if has_conv:
                    self.params += [md.skip.weight, md.skip.bias]
                self.zero_params.append(md.affine.weight)
            else:
                d.cin, d.cout, d.is_conv = md.in_channels, md.out_channels, 1
                self.params += [md.weight, md.bias]
            dst.append(d)
            self.block_keys.append(name)
            self.block_modules.append(md)
        self.params += [module.out_norm.weight, module.out_norm.bias, module.out_conv.weight, module.out_conv.bias]
        if module.map_label is not None:
            self.zero_params.append(module.map_label.weight)
        h = C.c_void_p()
        ea, da = (UNetBlockDesc * len(enc))(*enc), (UNetBlockDesc * len(dec))(*dec)
        check(lib().pub_unet_create(ea, len(enc), da, len(dec), self.in_ch, self.out_ch, C.c_float(module.dropout),
                                    dtype, C.byref(h)), "pub_unet_create")
        self.handle = h
        assert lib().pub_unet_num_params(h) == len(self.params)

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and _lib is not None:
                _lib.pub_unet_destroy(self.handle)
        except Exception:
            pass

    def forward(self, x, training, nhwc_out=False, seed=None):
        require_cuda(x, *self.params)
        if x.dim() != 4 or x.shape[1] != self.in_ch:
            raise ValueError(f"UNet expects [B,{self.in_ch},H,W], got {tuple(x.shape)}")
        if seed is None:
            seed = 0
            if training and self.module.dropout > 0:
                base, off = _Rng.next()
                seed = (base * 0x9E3779B97F4A7C15 + off + 1) & 0x7FFFFFFFFFFFFFFF
        self.last_seed = seed
        return _UNetFn.apply(self, x.float(), bool(training), bool(nhwc_out), int(seed), len(self.zero_params),
                             *self.params, *self.zero_params)

    def dropout_mask(self, block_key, B, H, W, seed):
        """Test hook: bool [B,C,h,w] keep-mask the engine uses in block `block_key` ("enc.<name>" /
        "dec.<name>") for `seed`."""
        idx = self.block_keys.index(block_key)
        md = self.block_modules[idx]
        h, w = H, W                      # output resolution of the block: walk the plan
        for m in self.block_modules[:idx + 1]:
            if getattr(m, "down", False):
                h, w = h // 2, w // 2
            if getattr(m, "up", False):
                h, w = h * 2, w * 2
        out = torch.empty(B, md.out_channels, h, w, device="cuda", dtype=torch.uint8)
        check(lib().pub_unet_dropout_mask(self.handle, idx, B, H, W, C.c_uint64(seed), ptr(out), stream()),
              "pub_unet_dropout_mask")
        return out.bool()


# ======================================================================================
# Gaussian encoder engine
# ======================================================================================
class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, target, *params):
        B, cx, H, W = x.shape
        L = lib()
        nbytes = L.pub_encoder_workspace_bytes(eng.handle, B, H, W)
        if nbytes == 0:
            raise NativeError("pub_encoder_workspace_bytes: " + L.pub_last_error().decode())
        ws = workspaces.take(nbytes, x.device)
        mu = torch.empty(B, eng.latent, device=x.device, dtype=torch.float32)
        sigma = torch.empty_like(mu)
        xc = x.contiguous()
        tc = target.contiguous() if target is not None else None
        check(L.pub_encoder_forward(eng.handle, B, H, W, ptr(xc), cx, ptr(tc), tc.shape[1] if tc is not None else 0,
                                    _ptr_table(params), ptr(mu), ptr(sigma), ptr(ws), C.c_size_t(nbytes), _backend,
                                    stream()), "pub_encoder_forward")
        ctx.eng, ctx.nbytes, ctx.shape = eng, nbytes, (B, H, W)
        if any(ctx.needs_input_grad):
            ctx.ws = ws
            ctx.save_for_backward(*params)
        else:
            ctx.ws = None
            workspaces.give(ws)
        return mu, sigma

    @staticmethod
    def backward(ctx, dmu, dsigma):
        eng, (B, H, W) = ctx.eng, ctx.shape
        params = ctx.saved_tensors
        flat, gviews = _flat_grads(list(params))
        dmu = (dmu if dmu is not None else torch.zeros(B, eng.latent, device=flat.device)).contiguous()
        dsigma = (dsigma if dsigma is not None else torch.zeros(B, eng.latent, device=flat.device)).contiguous()
        ws = _consume_ws(ctx, "AxisAlignedConvGaussian")
        check(lib().pub_encoder_backward(eng.handle, B, H, W, ptr(dmu), ptr(dsigma), _ptr_table(params),
                                         _ptr_table(gviews), ptr(ws), C.c_size_t(ctx.nbytes), _backend, stream()),
              "pub_encoder_backward")
        workspaces.give(ws)
        _notify(flat, list(params), gviews)
        return (None, None, None) + tuple(gviews)


class EncoderEngine:
    def __init__(self, module, dtype):
        import torch.nn as nn
        self.dtype, self.latent = dtype, module.latent_dim
        convs = [m for m in module.encoder if isinstance(m, nn.Conv2d)]
        self.params = []
        for c in convs:
            self.params += [c.weight, c.bias]
        self.params += [module.conv_mu.weight, module.conv_mu.bias, module.conv_log_sigma.weight, module.conv_log_sigma.bias]
        filt = (C.c_int32 * len(module.num_filters))(*module.num_filters)
        h = C.c_void_p()
        check(lib().pub_encoder_create(module.input_channels, filt, len(module.num_filters), module.latent_dim, dtype,
                                       C.byref(h)), "pub_encoder_create")
        self.handle = h
        assert lib().pub_encoder_num_params(h) == len(self.params)

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and _lib is not None:
                _lib.pub_encoder_destroy(self.handle)
        except Exception:
            pass

    def forward(self, x, target):
        require_cuda(x, target, *self.params)
        if x.requires_grad or (target is not None and target.requires_grad):
            raise NativeError("gradients w.r.t. the encoder inputs are not implemented (the reference never needs them)")
        return _EncoderFn.apply(self, x.float(), target.float() if target is not None else None, *self.params)


# ======================================================================================
# latent ops
# ======================================================================================
class _RsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, sigma, n, eps):
        B, L = mu.shape
        z = torch.empty(n, B, L, device=mu.device, dtype=torch.float32)
        eps_out = torch.empty_like(z)
        seed, off = _Rng.next()
        e = eps.contiguous().float() if eps is not None else None
        if e is not None and e.numel() != n * B * L:
            raise ValueError(f"eps must have {n}x{B}x{L} elements")
        check(lib().pub_rsample_forward(ptr(mu.contiguous()), ptr(sigma.contiguous()), ptr(e), C.c_uint64(seed),
                                        C.c_uint64(off), n, B, L, ptr(z), ptr(eps_out), stream()), "pub_rsample_forward")
        ctx.save_for_backward(eps_out)
        ctx.dims = (n, B, L)
        return z

    @staticmethod
    def backward(ctx, dz):
        (eps,) = ctx.saved_tensors
        n, B, L = ctx.dims
        dmu = torch.empty(B, L, device=dz.device, dtype=torch.float32)
        dsig = torch.empty_like(dmu)
        check(lib().pub_rsample_backward(ptr(dz.contiguous()), ptr(eps), n, B, L, ptr(dmu), ptr(dsig), stream()),
              "pub_rsample_backward")
        return dmu, dsig, None, None


def rsample(mu, sigma, n, eps=None):
    """z[n,B,L] = mu + sigma * eps, eps ~ N(0,1) from the engine's Philox stream (or injected)."""
    require_cuda(mu, sigma)
    return _RsampleFn.apply(mu, sigma, int(n), eps)


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mq, sq, mp, sp):
        B, L = mq.shape
        t = [a.contiguous().float() for a in (mq, sq, mp, sp)]
        kl = torch.empty(B, device=mq.device, dtype=torch.float32)
        check(lib().pub_kl_normal_forward(*[ptr(a) for a in t], B, L, ptr(kl), stream()), "pub_kl_normal_forward")
        ctx.save_for_backward(*t)
        return kl

    @staticmethod
    def backward(ctx, dkl):
        t = ctx.saved_tensors
        B, L = t[0].shape
        outs = [torch.empty_like(t[0]) if ctx.needs_input_grad[i] else None for i in range(4)]
        check(lib().pub_kl_normal_backward(ptr(dkl.contiguous()), *[ptr(a) for a in t], B, L, *[ptr(o) for o in outs],
                                           stream()), "pub_kl_normal_backward")
        return tuple(outs)


def kl_normal(mq, sq, mp, sp):
    """KL(N(mq,sq) || N(mp,sp)) summed over the latent axis -> [B]."""
    require_cuda(mq, sq, mp, sp)
    return _KLFn.apply(mq, sq, mp, sp)


# ======================================================================================
# fcomb
# ======================================================================================
def _fcomb_args(mod, feat, z, nhwc, out):
    a = FcombArgs()
    l0, l1, l2 = mod.layers[0], mod.layers[2], mod.layers[4]
    M, B, L = z.shape
    if nhwc:
        _, H, W, F = feat.shape
        a.feat_nchw, a.dtype = 0, (BF16 if feat.dtype == torch.bfloat16 else F32)
    else:
        _, F, H, W = feat.shape
        a.feat_nchw, a.dtype = 1, F32
        for i in range(4):
            a.stride[i] = feat.stride(i)
    a.feat, a.z = feat.data_ptr(), z.data_ptr()
    a.w0, a.b0, a.w1, a.b1, a.w2, a.b2 = (t.data_ptr() for t in (l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias))
    a.out = out.data_ptr() if out is not None else None
    a.B, a.H, a.W, a.F, a.L, a.C, a.M = B, H, W, F, L, l2.weight.shape[0], M
    return a


class _FcombFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, nhwc, feat, z, *params):
        M, B, L = z.shape
        H, W = (feat.shape[1], feat.shape[2]) if nhwc else (feat.shape[2], feat.shape[3])
        if feat.shape[0] != B:
            raise ValueError(f"feature batch {feat.shape[0]} != latent batch {B}")
        if nhwc:
            feat = feat.contiguous()
        elif feat.dtype != torch.float32:
            feat = feat.float()
        z = z.contiguous().float()
        out = torch.empty(B, M, mod.num_classes, H, W, device=z.device, dtype=torch.float32)
        a = _fcomb_args(mod, feat, z, nhwc, out)
        nws = lib().pub_fcomb_forward_workspace(C.byref(a))
        ws = torch.empty(nws, device=z.device, dtype=torch.uint8)
        check(lib().pub_fcomb_forward(C.byref(a), ptr(ws), C.c_size_t(nws), stream()), "pub_fcomb_forward")
        ctx.mod, ctx.nhwc = mod, nhwc
        ctx.save_for_backward(feat, z, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        feat, z, *params = ctx.saved_tensors
        mod, nhwc = ctx.mod, ctx.nhwc
        a = _fcomb_args(mod, feat, z, nhwc, None)
        nbytes = lib().pub_fcomb_backward_workspace(C.byref(a))
        ws = torch.empty(nbytes, device=z.device, dtype=torch.uint8)
        flat, g = _flat_grads(params)
        dz = torch.empty_like(z) if ctx.needs_input_grad[3] else None
        dfeat = None
        if ctx.needs_input_grad[2]:
            dfeat = torch.empty(feat.shape, device=feat.device, dtype=feat.dtype)   # contiguous, same logical layout
        check(lib().pub_fcomb_backward(C.byref(a), ptr(dout.contiguous()), ptr(dfeat), ptr(dz), ptr(g[0]), ptr(g[1]),
                                       ptr(g[2]), ptr(g[3]), ptr(g[4]), ptr(g[5]), ptr(ws), C.c_size_t(nbytes), stream()),
              "pub_fcomb_backward")
        _notify(flat, list(params), g)
        return (None, None, dfeat, dz) + tuple(g)


def fcomb_apply(mod, feat, z, nhwc=False):
    """feat: NHWC engine tensor (nhwc=True) or reference-layout [B,F,H,W] f32 (any strides);
    z [M,B,L] -> [B,M,C,H,W] f32."""
    require_cuda(feat, z, mod.layers[0].weight)
    l0, l1, l2 = mod.layers[0], mod.layers[2], mod.layers[4]
    return _FcombFn.apply(mod, bool(nhwc), feat, z, l0.weight, l0.bias, l1.weight, l1.bias, l2.weight, l2.bias)


# ======================================================================================
# losses
# ======================================================================================
def _loss_ws(B, Cc, HW, device):
    n = lib().pub_loss_workspace(B, Cc, HW)
    return torch.empty(n, device=device, dtype=torch.uint8), n


class _EnsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ens, target, kind, alpha):
        B, M, Cc, H, W = ens.shape
        ens, target = ens.contiguous().float(), target.contiguous().float()
        loss = torch.empty((), device=ens.device, dtype=torch.float32)
        dens = torch.empty_like(ens) if ctx.needs_input_grad[0] else None
        ws, n = _loss_ws(B, Cc, H * W, ens.device)
        check(lib().pub_ensemble_loss(ptr(ens), ptr(target), B, M, Cc, H * W, kind, C.c_float(alpha), ptr(loss),
                                      ptr(dens), ptr(ws), C.c_size_t(n), stream()), "pub_ensemble_loss")
        ctx.dens = dens
        return loss

    @staticmethod
    def backward(ctx, dloss):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        d, ctx.dens = ctx.dens, None
        if d is None:
            raise NativeError("ensemble_loss: second backward through the same loss node (its stored gradient is scaled "
                              "in place and released; retain_graph is not supported)")
        check(lib().pub_scale_by_device_scalar(ptr(d), ptr(dloss.contiguous().float()), C.c_int64(d.numel()), stream()),
              "pub_scale_by_device_scalar")
        return d, None, None, None


def ensemble_loss(ens, target, kind="afcrps", alpha=0.95):
    require_cuda(ens, target)
    if target.dim() != 4:
        raise NotImplementedError("per-member targets ([B,M,C,H,W]) are not used by the reference elbo")
    return _EnsLossFn.apply(ens, target, {"afcrps": 0, "crps": 1}[kind], float(alpha))


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, target):
        B, Cc, H, W = out.shape
        out, target = out.contiguous().float(), target.contiguous().float()
        res = torch.empty(1 + Cc, device=out.device, dtype=torch.float32)
        dout = torch.empty_like(out) if ctx.needs_input_grad[0] else None
        ws, n = _loss_ws(B, Cc, H * W, out.device)
        check(lib().pub_l1_loss(ptr(out), ptr(target), B, Cc, H * W, ptr(res), ptr(dout), ptr(ws), C.c_size_t(n),
                                stream()), "pub_l1_loss")
        ctx.dout = dout
        loss, per_var = res[0], res[1:]
        ctx.mark_non_differentiable(per_var)
        return loss, per_var

    @staticmethod
    def backward(ctx, dl, _dpv):
        if not ctx.needs_input_grad[0]:
            return None, None
        d, ctx.dout = ctx.dout, None
        if d is None:
            raise NativeError("l1_loss: second backward through the same loss node (retain_graph is not supported)")
        check(lib().pub_scale_by_device_scalar(ptr(d), ptr(dl.contiguous().float()), C.c_int64(d.numel()), stream()),
              "pub_scale_by_device_scalar")
        return d, None


def l1_loss(out, target):
    """-> (mean |out-target|, per-variable means [C])  (src/prob_unet.py:357-362)."""
    require_cuda(out, target)
    return _L1Fn.apply(out, target)


class _MsSsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, alpha, beta, lam):
        B, Cc, H, W = pred.shape
        pred, target = pred.contiguous().float(), target.contiguous().float()
        out3 = torch.empty(3, device=pred.device, dtype=torch.float32)
        dpred = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        n = lib().pub_msssim_workspace(B, Cc, H, W)
        if n == 0:
            raise NativeError(f"pub_msssim_workspace: {lib().pub_last_error().decode()}")
        ws = torch.empty(n, device=pred.device, dtype=torch.uint8)
        check(lib().pub_wmse_msssim_loss(ptr(pred), ptr(target), B, Cc, H, W, C.c_float(alpha), C.c_float(beta),
                                         C.c_float(lam), ptr(out3), ptr(dpred), ptr(ws), C.c_size_t(n), stream()),
              "pub_wmse_msssim_loss")
        ctx.dpred = dpred
        loss, wmse, ms = out3[0], out3[1], out3[2]
        ctx.mark_non_differentiable(wmse, ms)
        return loss, wmse, ms

    @staticmethod
    def backward(ctx, dl, _dw, _dm):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        d, ctx.dpred = ctx.dpred, None
        if d is None:
            raise NativeError("wmse_ms_ssim: second backward through the same loss node (retain_graph is not supported)")
        check(lib().pub_scale_by_device_scalar(ptr(d), ptr(dl.contiguous().float()), C.c_int64(d.numel()), stream()),
              "pub_scale_by_device_scalar")
        return d, None, None, None, None


def wmse_ms_ssim(pred, target, alpha, beta, lam, data_range=None):
    """-> (lam*WMSE + (1-lam)*(1-MS-SSIM), WMSE, 1-MS-SSIM) as 0-dim device tensors (src/prob_unet_utils.py:270-305).
    data_range is always inferred from the target on the device, as the reference's elbo does."""
    require_cuda(pred, target)
    if data_range is not None:
        raise NotImplementedError("an explicit data_range is never passed on the reference's hot path")
    return _MsSsimFn.apply(pred, target, float(alpha), float(beta), float(lam))


# ======================================================================================
# ensemble metrics
# ======================================================================================
def ensemble_metrics(preds, hr, lrinterp=None, std_hr=None):
    """preds [T,M,3,H,W], hr [T,3,H,W] -> (crps [T,3], mae [T,3]) on the device.  With lrinterp/std_hr the
    members are standardised residuals and are mapped to real units inside the kernel."""
    require_cuda(preds, hr)
    T, M, Cc, H, W = preds.shape
    crps = torch.empty(T, Cc, device=preds.device, dtype=torch.float32)
    mae = torch.empty_like(crps)
    tr = lrinterp is not None
    sh = std_hr.reshape(-1).contiguous().float() if tr else None
    li = lrinterp.contiguous().float() if tr else None
    nws = lib().pub_ensemble_metrics_workspace(T, Cc, H * W)
    ws = torch.empty(nws, device=preds.device, dtype=torch.uint8)
    check(lib().pub_ensemble_metrics(ptr(preds.contiguous().float()), ptr(hr.contiguous().float()), ptr(li), ptr(sh),
                                     int(tr), T, M, Cc, H * W, ptr(crps), ptr(mae), ptr(ws), C.c_size_t(nws), stream()),
          "pub_ensemble_metrics")
    return crps, mae


# ======================================================================================
# ensemble post-processing diagnostics (results.ipynb cells 4 and 15)
# ======================================================================================
_psd_tables = {}


def _psd_table(H, device):
    key = (H, str(device))
    if key not in _psd_tables:
        n = lib().pub_psd_table_ints(H)
        host = torch.empty(n, dtype=torch.int32)
        used = lib().pub_psd_build_table(H, C.c_void_p(host.data_ptr()))
        if used < 0:
            raise NativeError("pub_psd_build_table: " + lib().pub_last_error().decode())
        _psd_tables[key] = host.to(device)
    return _psd_tables[key]


def radial_psd(data, transfo=False, units=False):
    """data [N,C,H,H] f32 -> (psd per field [N,C,H/2], mean over N [C,H/2]); see pub_radial_psd."""
    require_cuda(data)
    Nn, Cc, H, W = data.shape
    if H != W:
        raise ValueError("radial_psd expects square fields")
    d = data.contiguous().float()
    per = torch.empty(Nn, Cc, H // 2, device=d.device, dtype=torch.float32)
    mean = torch.empty(Cc, H // 2, device=d.device, dtype=torch.float32)
    check(lib().pub_radial_psd(ptr(d), Nn, Cc, H, int(transfo), int(units), ptr(_psd_table(H, d.device)), ptr(per),
                               ptr(mean), stream()), "pub_radial_psd")
    return per, mean


def histogram(values, edges):
    """np.histogram(values, bins=edges) -> int64 counts [len(edges) - 1] on the device."""
    require_cuda(values)
    v = values.contiguous().float().reshape(-1)
    e = torch.as_tensor(edges, dtype=torch.float64).to(v.device).contiguous()
    counts = torch.zeros(e.numel() - 1, device=v.device, dtype=torch.int64)
    check(lib().pub_histogram(ptr(v), C.c_int64(v.numel()), ptr(e), e.numel() - 1, ptr(counts), stream()), "pub_histogram")
    return counts

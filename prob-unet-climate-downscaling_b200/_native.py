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

F32, BF16 = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05 = 0, 1, 2
_DTYPE_NAMES = {"fp32": F32, "float32": F32, "f32": F32, "bf16": BF16, "bfloat16": BF16}
_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16}

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
    "pub_last_error", "pub_version", "pub_conv2d_forward", "pub_pack_conv_weight", "pub_conv2d_wgrad_workspace",
    "pub_conv2d_wgrad", "pub_nchw_to_nhwc", "pub_nhwc_to_nchw", "pub_unet_create", "pub_unet_destroy",
    "pub_unet_num_params", "pub_unet_workspace_bytes", "pub_unet_forward", "pub_unet_backward",
    "pub_unet_dropout_mask", "pub_encoder_create", "pub_encoder_destroy", "pub_encoder_num_params",
    "pub_encoder_workspace_bytes", "pub_encoder_forward", "pub_encoder_backward", "pub_rsample_forward",
    "pub_rsample_backward", "pub_kl_normal_forward", "pub_kl_normal_backward", "pub_fcomb_forward",
    "pub_fcomb_backward_workspace", "pub_fcomb_backward", "pub_loss_workspace", "pub_ensemble_loss", "pub_l1_loss",
    "pub_scale_by_device_scalar", "pub_ensemble_metrics", "pub_adamw_step",
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
        for name in ("pub_conv2d_wgrad_workspace", "pub_unet_workspace_bytes", "pub_encoder_workspace_bytes",
                     "pub_fcomb_backward_workspace", "pub_loss_workspace"):
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


def conv2d_nhwc(x0, w_packed, bias=None, x1=None, res=None, mask=None, relu=False, ksize=3, backend=None, out=None):
    """x0/x1/res/mask: NHWC views [B,H,W,C] (last-dim-contiguous, arbitrary pixel stride)."""
    require_cuda(x0, w_packed)
    B, H, W, c0 = x0.shape
    dt = BF16 if x0.dtype == torch.bfloat16 else F32
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
    check(lib().pub_conv2d_forward(C.byref(a), stream()), "pub_conv2d_forward")
    return y


def conv2d_wgrad_nhwc(x0, dy, ksize, x1=None, want_bias=True, backend=None):
    B, H, W, c0 = x0.shape
    dt = BF16 if x0.dtype == torch.bfloat16 else F32
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


def conv2d(x, weight, bias):
    """Standalone NCHW f32 conv (networks.Conv2d.forward without resampling)."""
    dt = resolve_dtype(None)
    xh = nchw_to_nhwc(x, dt)
    y = conv2d_nhwc(xh, pack_conv_weight(weight, dt), bias, ksize=weight.shape[-1])
    return nhwc_to_nchw(y)

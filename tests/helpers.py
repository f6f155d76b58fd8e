"""Shared test helpers (also used by tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "prob-unet-climate-downscaling_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def dezero(model, seed=43):
    """SURVEY.md 8c oracle recipe step 2: conv1 / out_conv weights are zero at init, which makes
    unet(x) == 0 and parity vacuous -> re-randomise all-zero weight tensors and perturb the FiLM biases."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.dim() > 1 and float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        for n, p in model.named_parameters():
            if n.endswith("affine.bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.1)


def canonical_model(latent_dim=32, loss_type="afcrps", compute_dtype=None, beta=(1.0, 1.0, 0.0), device=None):
    """Our ProbabilisticUNet with the reference's seed-42 init + de-zero (built on CPU, optionally moved)."""
    import prob_unet
    saved = prob_unet.device
    prob_unet.device = torch.device("cpu")          # construct on CPU so that RNG order == reference
    try:
        torch.manual_seed(42)
        m = prob_unet.ProbabilisticUNet(3, 3, latent_dim, [32, 64, 128, 256], 32, [1, 2, 4, 8], *beta,
                                        loss_type=loss_type, compute_dtype=compute_dtype)
    finally:
        prob_unet.device = saved
    dezero(m)
    m.eval()
    return m.to(device) if device is not None else m


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grad_projection(g):
    """(g . r1, g . r2) in float64 with r1[k] = sin(1 + 0.7311 k), r2[k] = cos(0.1 + 2.399963 k) over the flattened
    (reference OIHW / row-major) tensor: closed-form, box-independent signed checksums of a gradient."""
    v = torch.as_tensor(g).detach().double().cpu().reshape(-1).numpy()
    k = np.arange(v.size, dtype=np.float64)
    return [float(np.dot(v, np.sin(1.0 + 0.7311 * k))), float(np.dot(v, np.cos(0.1 + 2.399963 * k)))]


def proj_close(got, ref, norm, tol):
    """|projection error| <= tol * ||g|| * sqrt(n/2)-free bound: a projection of a vector with relative error e has
    absolute error <= e * ||g|| * ||r||/sqrt(n) on average; we compare against tol * ||g||."""
    return abs(got[0] - ref[0]) <= tol * norm + 1e-12 and abs(got[1] - ref[1]) <= tol * norm + 1e-12


def unpack_masks(golden, topology_keys):
    """Dropout keep-masks stored by make_golden (packed bits, reference call order) -> {block key: bool tensor}."""
    shapes = golden["A_drop_maskshapes"]
    bits = np.unpackbits(golden["A_drop_maskbits"])
    out, off = {}, 0
    keys = [k for k in topology_keys]
    assert len(keys) == len(shapes)
    for k, shp in zip(keys, shapes):
        n = int(np.prod(shp))
        npad = (n + 7) // 8 * 8
        out[k] = torch.from_numpy(bits[off:off + n].astype(bool).reshape(tuple(int(s) for s in shp)))
        off += npad
    return out

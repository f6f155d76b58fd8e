"""the reference's OWN train / eval / sample loops (src/train_prob_unet_model.py:105-256), imported
unmodified with our host mirror first on sys.path (the import-shadowing recipe of INTEGRATION.md section 1), drive our
ProbabilisticUNet: `model.elbo(inputs, targets, timestamps, M=...)` must unpack into 3 values, `recon_list[0]` must be
a number np.mean accepts, `kl_div.mean().item()` and `loss.backward()` + `torch.optim.AdamW.step()` must work, and
`model(inputs, t=timestamps, training=False)` must bind.

CPU only and only where /root/reference exists (the authoring container; it does not travel to the GPU box, and no
reference source is copied into this repo).  There is no CPU product path, so for THIS host-logic test the three
native entry points the module mirror calls are replaced by oracle-backed stand-ins (tests may use the oracle as a
fake backend); everything above them -- our prob_unet.py / networks.py mirror -- is the real code under test.  The
kernels themselves are covered by the `-m gpu` parity tests."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from helpers import PKG, canonical_model
from oracle import probunet_oracle as O

REF = os.environ.get("PROBUNET_REFERENCE", "/root/reference") + "/src"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_prob_unet_model.py")),
                                reason="the reference tree is only present in the authoring container")


@pytest.fixture()
def reference_train_module():
    """import train_prob_unet_model from the reference with matplotlib / the data stack stubbed and OUR flat modules
    (prob_unet, networks, ...) ahead of the reference's on sys.path."""
    saved_path = list(sys.path)
    stubs = ("matplotlib", "matplotlib.pyplot", "climex_utils")
    for name in stubs:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    sys.path.insert(0, PKG)                      # the mirror shadows the reference's prob_unet / networks
    sys.modules.pop("train_prob_unet_model", None)
    import train_prob_unet_model as tm
    import prob_unet
    assert os.path.dirname(os.path.abspath(prob_unet.__file__)) == os.path.abspath(PKG)
    assert os.path.abspath(tm.__file__).startswith(os.path.abspath(REF))
    yield tm
    sys.path[:] = saved_path
    for k in stubs + ("train_prob_unet_model",):
        sys.modules.pop(k, None)


class _FakeBackend:
    """Oracle-backed stand-ins for the native calls prob_unet.py makes (CPU, autograd through the real Parameters)."""

    def __init__(self, model, monkeypatch):
        import _native
        import networks
        import prob_unet
        self.m = model
        cfg = O.ProbUNetCfg(latent_dim=model.latent_dim)
        sd = lambda: {k: v for k, v in model.state_dict(keep_vars=True).items()}   # noqa: E731

        def unet_forward(net, x, noise_labels=None, class_labels=None, augment_labels=None, _nhwc_out=False):
            masks = None
            if net.training:                         # train(): Bernoulli dropout, as F.dropout in the reference
                enc, dec = O.unet_topology(cfg.unet())
                masks, h = {}, x.shape[-1]
                for b in enc + dec:
                    if b.is_conv:
                        continue
                    h = h * 2 if b.up else (h // 2 if b.down else h)
                    masks[b.key] = torch.rand(x.shape[0], b.cout, h, h) >= 0.1
            return O.unet_forward(sd(), x, cfg.unet(), drop_masks=masks)

        def encoder_forward(enc, x, target=None):
            name = "posterior" if enc.posterior else "prior"
            mu, sig = O.gaussian_encoder(sd(), name, x, target if enc.posterior else None, cfg.num_filters)
            return prob_unet.LatentGaussian(mu, sig)

        def fcomb_apply(mod, feat, z, nhwc=False):
            return torch.stack([O.fcomb(sd(), feat, z[i]) for i in range(z.shape[0])], dim=1)

        def rsample(mu, sigma, n, eps=None):
            e = eps if eps is not None else torch.randn(n, *mu.shape)
            return mu + sigma * e.reshape(n, *mu.shape)

        monkeypatch.setattr(networks.UNet, "forward", unet_forward)
        monkeypatch.setattr(prob_unet.AxisAlignedConvGaussian, "forward", encoder_forward)
        monkeypatch.setattr(_native, "fcomb_apply", fcomb_apply)
        monkeypatch.setattr(_native, "rsample", rsample)
        monkeypatch.setattr(_native, "kl_normal", O.kl_normal)
        monkeypatch.setattr(_native, "l1_loss", lambda out, t: ((out - t).abs().mean(), (out - t).abs().mean(dim=(0, 2, 3))))
        monkeypatch.setattr(_native, "ensemble_loss", lambda ens, t, kind="afcrps", alpha=0.95:
                            O.afcrps_loss(ens, t, alpha) if kind == "afcrps" else O.crps_loss(ens, t))


class _Dataset:
    """What the loops touch of climex2torch: residual_to_hr (src/climex_utils.py:284-285) and the plot hook."""
    std_hr = torch.ones(3, 1, 1)

    def residual_to_hr(self, residual, lrinterp):
        return lrinterp + residual * (self.std_hr + 1e-10)

    def plot_sample_batch(self, *a, **k):
        return "fig", "axs"


def _loader(n_batches, B, res):
    from climex_synth import make_fields
    batches = []
    for i in range(n_batches):
        f = make_fields(B, res, res, 8, seed=50 + i)
        f["timestamps"] = torch.arange(B).float()
        f["timestamps_float"] = torch.arange(B).float()
        batches.append(f)

    class L(list):
        dataset = _Dataset()
    return L(batches)


@pytest.mark.parametrize("scalars", [True, "lazy"])
def test_reference_train_eval_and_sample_loops_drive_the_mirror(reference_train_module, monkeypatch, scalars):
    """scalars="lazy": elbo returns float-like LazyScalars instead of .item() floats (no host sync between forward and
    backward); the reference's loops -- list.append + np.mean per epoch -- must not notice."""
    tm = reference_train_module
    torch.set_num_threads(8)
    model = canonical_model(latent_dim=16)                      # afcrps is the default loss_type (src/main.py:1,136)
    model.sync_scalars = scalars
    _FakeBackend(model, monkeypatch)
    loader = _loader(2, 2, 32)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)        # src/main.py:103
    before = model.fcomb.layers[4].bias.detach().clone()
    crps, kl = tm.train_probunet_step(model, loader, opt, 0, 1, torch.device("cpu"), ensemble_size=3)
    assert np.isfinite(crps) and np.isfinite(kl) and crps > 0 and kl > 0
    assert model.training and not torch.equal(before, model.fcomb.layers[4].bias.detach())   # a real update happened
    # elbo touches every parameter (the zero-input map_label / affine weights get their exact-zero gradients from the
    # native engine only -- tests/test_gpu_model.py -- not from this fake backend)
    assert all(p.grad is not None for n, p in model.named_parameters()
               if not n.endswith("affine.weight") and not n.endswith("map_label.weight"))
    ec, ek = tm.eval_probunet_model(model, loader, torch.device("cpu"), ensemble_size=2)
    assert np.isfinite(ec) and np.isfinite(ek) and not model.training
    preds, (fig, axs) = tm.sample_probunet_model(model, loader, 0, torch.device("cpu"))
    assert preds.shape == (2, 3, 3, 32, 32) and fig == "fig"
    assert model.prior_latent_space is not None


def test_elbo_return_arity_follows_the_loss_type_like_the_three_reference_variants(monkeypatch):
    """src/prob_unet.py:267 (5 values), :317 (3 values), :381 (4 values)."""
    import inspect
    import prob_unet
    sig = inspect.signature(prob_unet.ProbabilisticUNet.elbo)
    sig.bind(None, "inputs", "targets", "timestamps", M=5)                        # train_prob_unet_model.py:133-136
    inspect.signature(prob_unet.ProbabilisticUNet.forward).bind(None, "inputs", t="ts", training=False)   # :245
    inspect.signature(prob_unet.ProbabilisticUNet.forward).bind(None, "inputs", target=None, t="ts", training=False)
    model = canonical_model(latent_dim=16)
    _FakeBackend(model, monkeypatch)
    f = _loader(1, 2, 32)[0]
    model.loss_type = "l1"
    out = model.elbo(f["inputs"], f["targets"], None)
    assert len(out) == 4 and len(out[1]) == 3 and all(isinstance(v, float) for v in out[1]) and out[3].shape == (2,)
    model.loss_type = "crps"
    out = model.elbo(f["inputs"], f["targets"], None, M=2)
    assert len(out) == 3 and isinstance(out[1][0], float) and out[2].shape == (2,)
    with pytest.raises(ValueError):
        model.elbo(f["inputs"], f["targets"], None, M=1)                          # src/prob_unet.py:282-283
    # the lazy policy returns the same numbers as float-likes
    from lazy_scalar import LazyScalar
    torch.manual_seed(3)
    strict = model.elbo(f["inputs"], f["targets"], None, M=2)[1][0]
    model.sync_scalars = "lazy"
    torch.manual_seed(3)
    lazy = model.elbo(f["inputs"], f["targets"], None, M=2)[1][0]
    assert isinstance(lazy, LazyScalar) and float(lazy) == strict
    model.loss_type = "l1"
    out = model.elbo(f["inputs"], f["targets"], None)
    assert len(out[1]) == 3 and all(isinstance(v, LazyScalar) for v in out[1]) and np.isfinite(np.mean(out[1]))

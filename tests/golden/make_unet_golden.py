"""Golden vectors for BASELINE configs[1] (deterministic U-Net) from the REAL reference (authoring container only).

    python tests/golden/make_unet_golden.py

Imports /root/reference/src/networks.py unmodified and builds the architecture of src/deterministic_unet_main.py:52
(`UNet(img_resolution, in_channels=3, out_channels=3, ...)`, model_channels = 16 and channel_mult = [1,4,8,16] at their
defaults) with manual_seed(42).  The driver passes label_dim=0, but in this snapshot that path cannot run:
`UNet.forward` then feeds a [B,1] zero embedding into every block's `affine` Linear(64 -> C) and raises a shape
error (src/networks.py:315-316, :173).  The golden is therefore generated with label_dim=1 (the class default): the
embedding is `map_label(zeros) = 0`, i.e. exactly the zero conditioning the label_dim=0 branch intends, on the same
16/64/128/256-channel network.  Then de-zeroes the zero-initialised tensors (tests/helpers.dezero, seed 43), runs
forward + MSE + backward (src/trainmodel.py:158-160) on a seeded batch and stores inputs, output, loss, state_dict
checksums and every gradient norm in unet_golden.npz beside this script.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PROBUNET_REFERENCE", "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import dezero  # noqa: E402


def main():
    sys.path.insert(0, REF)
    import networks as ref_networks
    sys.path.pop(0)
    assert os.path.abspath(ref_networks.__file__).startswith(os.path.abspath(REF)), ref_networks.__file__
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(42)
    net = ref_networks.UNet(img_resolution=(64, 64), in_channels=3, out_channels=3, label_dim=1, use_diffuse=False)
    out = {"model_channels": np.array(net.enc["64x64_conv"].out_channels)}
    sd0 = net.state_dict()
    out["sd_keys"] = np.array(list(sd0.keys()))
    out["sd_sum"] = np.array([float(v.double().sum()) for v in sd0.values()])
    dezero(net)
    sd1 = net.state_dict()
    out["sd1_sum"] = np.array([float(v.double().sum()) for v in sd1.values()])
    out["sd1_abssum"] = np.array([float(v.double().abs().sum()) for v in sd1.values()])
    net.eval()
    g = torch.Generator().manual_seed(6)
    x, y = torch.randn(2, 3, 64, 64, generator=g), torch.randn(2, 3, 64, 64, generator=g)
    out["x"], out["y"] = x.numpy(), y.numpy()
    pred = net(x)          # UNet.forward(self, x): src/networks.py:299 (trainmodel.py passes class_labels=, which the snapshot rejects)
    loss = torch.nn.functional.mse_loss(pred, y)
    loss.backward()
    out["pred"] = pred.detach().numpy()
    out["loss"] = np.array(float(loss))
    names, norms = [], []
    for n, p in net.named_parameters():
        names.append(n)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    out["grad_names"], out["grad_norm"] = np.array(names), np.array(norms)
    path = os.path.join(HERE, "unet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), "model_channels", int(out["model_channels"]))


if __name__ == "__main__":
    main()

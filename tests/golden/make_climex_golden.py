"""Golden vectors of the reference's dataset transform, produced by the REAL climex2torch class
(/root/reference/src/climex_utils.py) in this container: its module-level imports (dask, xarray, cartopy,
matplotlib ...) are stubbed, the instance is created without __init__ (which would read NetCDF files) and given a
synthetic `hr` tensor; __getitem__ / compute_stats / residual_to_hr then run unmodified.

    python tests/golden/make_climex_golden.py        # writes tests/golden/climex_golden.npz
"""
import os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("dask", "dask.distributed", "xarray", "bottleneck", "cftime", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
             "cartopy", "cartopy.crs"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["dask.distributed"].Client = object
sys.modules["cartopy"].crs = sys.modules["cartopy.crs"]
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
sys.path.insert(0, os.environ.get("PROBUNET_REFERENCE", "/root/reference") + "/src")
import climex_utils as CU  # noqa: E402

T, s, H = 10, 16, 64
g = torch.Generator().manual_seed(2024)
hr = torch.randn(T, 3, H, H, generator=g) * torch.tensor([2.0, 8.0, 9.0]).view(1, 3, 1, 1) + torch.tensor([1.0, 270.0, 280.0]).view(1, 3, 1, 1)
ds = object.__new__(CU.climex2torch)
ds.hr, ds.lowres_scale, ds.epsilon, ds.type, ds.lrstats = hr, s, 1e-10, "lrinterp_to_residuals", None
ds.timestamps = list(range(T)); ds.timestamps_float = [float(t) for t in range(T)]
items = [ds[i] for i in (0, 3, 7)]
out = {"hr": hr.numpy(), "scale": np.array(s), "idx": np.array([0, 3, 7]),
       "mean_lr": ds.lrstats[0][0].numpy(), "std_lr": ds.lrstats[0][1].numpy(),
       "mean_hr": ds.lrstats[1][0].numpy(), "std_hr": ds.lrstats[1][1].numpy()}
for k in ("inputs", "targets", "lr", "lrinterp"):
    out[k] = torch.stack([it[k] for it in items]).numpy()
res = torch.randn(3, 3, H, H, generator=g)
out["residual"] = res.numpy()
out["residual_to_hr"] = torch.stack([ds.residual_to_hr(res[i], items[i]["lrinterp"]) for i in range(3)]).numpy()
np.savez_compressed(os.path.join(HERE, "climex_golden.npz"), **out)
print("wrote", os.path.join(HERE, "climex_golden.npz"), {k: v.shape for k, v in out.items()})

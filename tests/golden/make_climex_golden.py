"""Golden vectors of the reference's dataset transform, produced by the REAL climex2torch class
(/root/reference/src/climex_utils.py) in this container: its module-level imports (dask, xarray, cartopy,
matplotlib ...) are stubbed, the instance is created without __init__ (which would read NetCDF files) and given a
synthetic `hr` tensor; __getitem__ / compute_stats / residual_to_hr then run unmodified.  Also pinned here: the
inverse variable transforms softplus / KToC / kgm2sTommday (src/climex_utils.py:32-50), their composition
invert_transfo_3vars (results.ipynb cell 2, executed from the notebook's own source) and metrics.compute_mae
(src/metrics.py:48-71).

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

# ---- inverse variable transforms (src/climex_utils.py:32-50) and their composition in results.ipynb cell 2
# (invert_transfo_3vars, the function cell 11 maps over every (t, member) before metrics.crps_over_groundtruth)
xs = torch.cat([torch.randn(4000, generator=g) * 6.0, torch.tensor([19.9999, 20.0, 20.0001, 25.0, 80.0, -30.0, 0.0])])
out["tf_x"] = xs.numpy()
out["tf_softplus"] = CU.softplus(xs.clone()).numpy()               # the real functions work in place: clone
out["tf_softplus_c0"] = CU.softplus(xs.clone(), c=0).numpy()
out["tf_ktoc"] = CU.KToC(xs.clone()).numpy()
out["tf_mmday"] = CU.kgm2sTommday(xs.clone()).numpy()
import json  # noqa: E402
nb = json.load(open(os.environ.get("PROBUNET_REFERENCE", "/root/reference") + "/src/notebooks/results.ipynb"))
cell2 = "".join(nb["cells"][2]["source"])
assert "def invert_transfo_3vars" in cell2
ns = {"torch": torch}
exec(compile(cell2, "results.ipynb:cell2", "exec"), ns)
stored = torch.randn(5, 3, 16, 16, generator=g) * torch.tensor([3.0, 9.0, 4.0]).view(1, 3, 1, 1) + torch.tensor([-1.0, 272.0, 3.0]).view(1, 3, 1, 1)
out["tf_stored"] = stored.numpy()
out["tf_real"] = ns["invert_transfo_3vars"](stored.clone()).numpy()

# ---- the per-pixel conversions of test_return_levels.ipynb cell 2 (composed from the real functions exactly as there:
# tasmax uses softplus(..., c=0) in THAT notebook)
hp = torch.randn(64, 3, 4, 4, generator=g) * torch.tensor([3.0, 9.0, 4.0]).view(1, 3, 1, 1) + torch.tensor([-1.0, 272.0, 3.0]).view(1, 3, 1, 1)
out["rl_hr"] = hp.numpy()
out["rl_pr"] = CU.kgm2sTommday(CU.softplus(hp[:, 0].clone())).numpy()
out["rl_tasmax"] = CU.KToC(hp[:, 1] + CU.softplus(hp[:, 2].clone(), c=0)).numpy()
out["rl_tasmin"] = CU.KToC(hp[:, 1].clone()).numpy()

# ---- results.ipynb cell 4 (psd / compute_psd_tensor), executed from the notebook's own source with the real scipy
cell4 = "".join(nb["cells"][4]["source"])
assert "def compute_psd_tensor" in cell4 and "def psd" in cell4
import scipy.stats as _stats  # noqa: E402
ns4 = {"torch": torch, "np": np, "stats": _stats, "cu": CU}
exec(compile(cell4, "results.ipynb:cell4", "exec"), ns4)
pf = torch.randn(6, 3, 32, 32, generator=g) * torch.tensor([2.0, 6.0, 3.0]).view(1, 3, 1, 1) + torch.tensor([0.5, 275.0, 4.0]).view(1, 3, 1, 1)
pf = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(pf, (1, 1, 1, 1), mode="circular"), 3, 1)   # some spectral slope
out["psd_fields"] = pf.numpy()
k4, p4 = ns4["psd"](pf[0, 1].clone())
out["psd_k"], out["psd_single"] = np.asarray(k4), np.asarray(p4)
for tf in (True, False):
    r4 = ns4["compute_psd_tensor"](pf.clone(), transfo=tf)
    out["psd_tensor_" + ("transfo" if tf else "plain")] = np.stack([r4[v] for v in ("pr", "tasmin", "tasmax")], axis=0)
r5 = ns4["compute_psd_tensor"](pf.clone().reshape(2, 3, 3, 32, 32), transfo=True)       # the [T, M, 3, H, W] form
out["psd_tensor_5d"] = np.stack([r5[v] for v in ("pr", "tasmin", "tasmax")], axis=0)
# ---- cell 15: np.histogram over np.linspace(min, max, NBINS) edges
hv = (torch.randn(5000, generator=g) * 3 + 1).numpy()
he = np.linspace(hv.min(), hv.max(), 100)
out["hist_values"], out["hist_edges"], out["hist_counts"] = hv, he, np.histogram(hv, bins=he)[0]

# ---- metrics.compute_mae (src/metrics.py:48-71) run for real; only the module-level `import pysteps` is stubbed
# (compute_mae never calls it).  crps_over_groundtruth needs pysteps itself and stays a restatement (oracle header).
sys.modules["pysteps"] = types.ModuleType("pysteps")
import metrics as RM  # noqa: E402
gt = torch.randn(4, 3, 12, 12, generator=g) * 5
pe = gt.unsqueeze(1) + torch.randn(4, 7, 3, 12, 12, generator=g)
mm, ma = RM.compute_mae(gt, pe)
md, mda = RM.compute_mae(gt, pe[:, 0])
out["mae_gt"], out["mae_pred"] = gt.numpy(), pe.numpy()
out["mae_ens"] = np.stack([ma[v] for v in ("pr", "tasmin", "tasmax")], axis=1)
out["mae_det"] = np.stack([mda[v] for v in ("pr", "tasmin", "tasmax")], axis=1)
out["mae_ens_means"] = np.array([mm[v] for v in ("pr", "tasmin", "tasmax")])
np.savez_compressed(os.path.join(HERE, "climex_golden.npz"), **out)
print("wrote", os.path.join(HERE, "climex_golden.npz"), {k: v.shape for k, v in out.items()})

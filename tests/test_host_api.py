"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror keeps the
reference's module surface, there is no CPU fallback, and the multi-rank host logic works over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import ROOT, PKG, canonical_model


def test_library_exports_every_declared_symbol():
    import _native as N
    so = os.path.join(PKG, "libprobunet_b200.so")
    if not os.path.exists(so):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, "include", "probunet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pub_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/probunet_b200.h but not exported"
    assert sorted(N.EXPORTS) == declared
    assert lib.pub_version() >= 100


def test_module_surface_matches_reference():
    m = canonical_model(latent_dim=16)
    assert m.latent_dim == 16 and m.beta_0 == 1.0
    m.beta_1 = 1e-4                                   # assignable each epoch (src/main.py:122-123)
    assert m.fcomb.layers[0].weight.shape == (32, 48, 1, 1) and m.fcomb.layers[0].in_channels == 48
    assert list(m.unet.enc.keys())[:4] == ["128x128_conv", "128x128_block0", "128x128_block1", "64x64_down"]
    assert list(m.unet.dec.keys())[0] == "16x16_in0" and len(m.unet.dec) == 17
    assert "unet.enc.64x64_down.conv0.resample_filter" in m.state_dict()
    assert m.prior.encoder[7].weight.shape == (64, 32, 3, 3) and m.posterior.encoder[0].weight.shape[1] == 6
    a = torch.arange(6.).reshape(2, 3)
    assert torch.equal(m.fcomb.tile(a, 1, 2), torch.repeat_interleave(a, 2, dim=1))
    # state_dict round trip
    m2 = canonical_model(latent_dim=16)
    m2.load_state_dict(m.state_dict())


def test_no_cpu_fallback():
    import _native as N
    m = canonical_model()
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(N.NativeError):
        m.unet(x)
    with pytest.raises(N.NativeError):
        m.prior(x)
    with pytest.raises(N.NativeError):
        m.fcomb(torch.zeros(1, 32, 64, 64), torch.zeros(1, 32))
    with pytest.raises(N.NativeError):
        m(x, x)


def test_shard_range_partitions():
    from parallel import shard_range
    for n in (0, 1, 7, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_workspace_pool_reuses_the_best_fit_and_stays_bounded():
    import _native
    pool = _native._WorkspacePool()
    dev = torch.device("cpu")
    big, small = pool.take(1000, dev), pool.take(100, dev)
    pool.give(big); pool.give(small)
    assert pool.take(90, dev) is small                 # ragged last batch: the smallest buffer that is large enough
    assert pool.take(500, dev) is big                  # a smaller request reuses the big block instead of pinning another
    assert pool.take(2000, dev).numel() == 2000        # nothing fits: allocate
    for n in range(10):
        pool.give(torch.empty(10 + n, dtype=torch.uint8))
    assert len(pool.free[dev]) == pool.MAX_FREE and min(b.numel() for b in pool.free[dev]) == 14   # CPU: keyed by device
    pool.clear()
    assert not pool.free


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from parallel import GradSynchronizer, gather_scores, shard_range
import _native
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
sync = GradSynchronizer().install()
flat = torch.full((1000,), float(r + 1))
_native._notify(flat)                      # what every engine backward calls with its flat gradient buffer
flat2 = torch.full((10,), float(10 * (r + 1)))
_native._notify(flat2)                     # a second sub-network: both collectives stay in flight ...
assert len(sync.pending) == 2
sync.wait_all()                            # ... until the optimizer (FusedAdamW.step) asks for the gradients
assert not sync.pending
assert torch.allclose(flat, torch.full((1000,), 3.0)), flat[:4]
assert torch.allclose(flat2, torch.full((10,), 30.0)), flat2[:4]
assert sync.calls == 2 and sync.bytes == 4040 and sync.world == 2
# the reduced buffer must BE the parameters' .grad (what autograd does with zero_grad(set_to_none=True)) ...
p1, p2 = torch.nn.Parameter(torch.zeros(6)), torch.nn.Parameter(torch.zeros(2, 2))
flat3 = torch.arange(10.) * (r + 1)
views = [flat3[:6].view(6), flat3[6:].view(2, 2)]
p1.grad, p2.grad = views
_native._notify(flat3, [p1, p2], views)
sync.wait_all()
assert torch.equal(p1.grad, torch.arange(6.) * 3) and torch.equal(p2.grad.reshape(-1), torch.arange(6., 10.) * 3)
# ... averaged for optimizers that do not fold 1/world into their update
flat4 = torch.arange(10.) * (r + 1)
views = [flat4[:6].view(6), flat4[6:].view(2, 2)]
p1.grad, p2.grad = views
_native._notify(flat4, [p1, p2], views)
sync.wait_all(average=True)
assert torch.equal(p1.grad, torch.arange(6.) * 1.5)
# ... a .grad autograd CLONED from the view (it was None before the backward) is overwritten with the reduced slice
flat7 = torch.arange(10.) * (r + 1)
views = [flat7[:6].view(6), flat7[6:].view(2, 2)]
p1.grad = p2.grad = None
_native._notify(flat7, [p1, p2], views)
p1.grad, p2.grad = views[0].clone(), views[1]
sync.wait_all()
assert torch.equal(p1.grad, torch.arange(6.) * 3) and p1.grad.data_ptr() != flat7.data_ptr()
# ... and an accumulated .grad (it existed before the backward) is an error, not a silent divergence of the ranks
flat5 = torch.ones(10)
views = [flat5[:6].view(6), flat5[6:].view(2, 2)]
p1.grad, p2.grad = views[0].clone(), views[1]
_native._notify(flat5, [p1, p2], views)
try:
    sync.wait_all()
    raise SystemExit("non-aliasing .grad was not detected")
except RuntimeError as ex:
    assert "set_to_none=True" in str(ex)
# a step whose collectives were never consumed (optimizer unaware of the synchronizer) is detected at the next backward
flat6 = torch.ones(10); views = [flat6[:6].view(6), flat6[6:].view(2, 2)]
_native._notify(flat6, [p1, p2], views)
try:
    _native._notify(torch.ones(10), [p1, p2], views)
    raise SystemExit("unconsumed step was not detected")
except RuntimeError as ex:
    assert "wait_all" in str(ex)
assert not sync.pending
b, e = shard_range(5, r, 2)
local = torch.arange(b, e).float().reshape(-1, 1).repeat(1, 3)
allv = gather_scores(local, [3, 2])
assert allv.shape == (5, 3) and torch.equal(allv[:, 0], torch.arange(5.)), allv
GradSynchronizer.uninstall()
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_two_rank_gloo_gradient_sync_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), PKG, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_custom_ops_are_registered_with_fake_kernels():
    """torch.ops.probunet_b200.* (custom_ops.py): schemas exist and shape propagation works without a GPU."""
    import custom_ops
    for name in custom_ops.OPS:
        assert hasattr(torch.ops.probunet_b200, name), name
    x = torch.empty(2, 16, 16, 32, device="meta", dtype=torch.bfloat16)
    w = torch.empty(64, 32, 3, 3, device="meta")
    assert torch.ops.probunet_b200.conv2d_nhwc(x, w, None, False).shape == (2, 16, 16, 64)
    feat, z = torch.empty(2, 32, 16, 16, device="meta"), torch.empty(5, 2, 8, device="meta")
    ws = [torch.empty(s, device="meta") for s in ((32, 40, 1, 1), (32,), (32, 32, 1, 1), (32,), (3, 32, 1, 1), (3,))]
    assert torch.ops.probunet_b200.fcomb(feat, z, *ws).shape == (2, 5, 3, 16, 16)
    loss, dens = torch.ops.probunet_b200.ensemble_loss(torch.empty(2, 5, 3, 8, 8, device="meta"),
                                                       torch.empty(2, 3, 8, 8, device="meta"), 0, 0.95)
    assert loss.shape == () and dens.shape == (2, 5, 3, 8, 8)
    with pytest.raises(Exception):                       # no CPU kernel is registered
        torch.ops.probunet_b200.kl_normal(*[torch.zeros(2, 4) for _ in range(4)])


def test_every_pdl_launched_kernel_waits_for_its_grid_dependency():
    """Source-level guard: a kernel launched through launch_pdl() may start before its predecessor has finished, so
    its body must reach pdl_enter() / pdl_wait() before touching global memory.  Every kernel name passed to
    launch_pdl must therefore contain one of the two calls (the placement is reviewed by hand; its absence is a bug)."""
    src = {}
    csrc = os.path.join(PKG, "csrc")
    for fn in os.listdir(csrc):
        if fn.endswith((".cu", ".cuh")):
            src[fn] = open(os.path.join(csrc, fn)).read()
    launched = set()
    for text in src.values():
        launched |= set(re.findall(r"launch_pdl\(\s*([A-Za-z_][A-Za-z0-9_]*)", text))
    launched -= {"void", "kernel"}                       # the helper's own declaration
    assert len(launched) >= 20, launched
    for name in sorted(launched):
        bodies = []
        for text in src.values():
            for m in re.finditer(r"__global__[^;{]*\b" + name + r"\s*\(", text):
                start = text.index("{", m.end())
                bodies.append(text[start:start + 6000])
        assert bodies, f"{name}: launched with launch_pdl but no __global__ definition found"
        for body in bodies:
            assert re.search(r"\bpdl_(enter|wait)\(\)", body), f"{name} is launched with launch_pdl() but never waits"


def test_reference_arm_of_bench_prints_the_contract_line():
    """`bench.py --impl reference` runs the CPU oracle's training step on the host cores (no GPU, no CUDA library)
    and prints ONE JSON line with the keys the driver reads."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_samples_per_s" and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("probunet_train_128x128_b64pergpu_afcrps_M15")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-9 * max(1.0, d["value"])


def test_lazy_scalar_behaves_like_the_float_the_reference_returns():
    """lazy_scalar.LazyScalar stands in for the Python floats of the reference's elbo() (recon_list[0], per-variable L1
    list): conversions, arithmetic, comparisons, formatting, and numpy reductions over lists of them."""
    import numpy as np
    from lazy_scalar import LazyScalar, lazy_list
    vals = lazy_list(torch.tensor([1.5, 2.5, -3.0]))
    assert [float(v) for v in vals] == [1.5, 2.5, -3.0] and vals[0].item() == 1.5
    assert np.mean(vals) == pytest.approx(1.0 / 3) and np.asarray(vals).dtype == np.float64
    assert float(np.mean(vals + [1.0])) == 0.5 and sum(vals) == 1.0 and max(vals) == 2.5
    assert vals[0] + 1 == 2.5 and 2 * vals[1] == 5.0 and vals[1] / 2 == 1.25 and 1 - vals[0] == -0.5 and -vals[2] == 3.0
    assert vals[2] < 0 < vals[0] <= 1.5 and vals[0] == 1.5 and vals[0] != vals[1] and abs(vals[2]) == 3.0
    assert f"{vals[0]:.3f}" == "1.500" and str(vals[1]) == "2.5" and round(vals[1]) == 2 and int(vals[2]) == -3
    assert np.float64(2.0) * vals[0] == 3.0 and isinstance(vals[0] + vals[1], float)
    single = LazyScalar(torch.tensor(4.0))
    assert float(single) == 4.0 and bool(single) and hash(single) == hash(4.0)

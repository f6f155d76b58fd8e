#!/usr/bin/env python
"""Benchmark of the Prob U-Net training hot path (BASELINE.json metric: train samples/s, 128x128
ClimEx-shaped grid, 1/2/4/8 B200) + ensemble members/s as an auxiliary figure.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: CPU oracle on the host cores

A "step" = zero_grad + ELBO forward (U-Net + prior + posterior + M x fcomb + afCRPS + KL) + backward +
AdamW update on one batch of synthetic ClimEx-shaped fields (+ gradient all-reduce when N > 1).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definitions of every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "prob-unet-climate-downscaling_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# algorithmic work per unit (SURVEY.md 8d / BASELINE.md section 3), canonical Prob U-Net, 128^2, L=32
GFLOP_TRAIN_SAMPLE_M1 = 95.25      # fwd + dgrad + wgrad, one ELBO member
GFLOP_PER_EXTRA_MEMBER = 0.311


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "ensemble", "detunet", "latent256"],
                    help="train: BASELINE.json configs[2] (train samples/s, the headline metric); ensemble: configs[3] "
                         "(prior-sampling ensemble members/s: M members per field + CRPS/MAE, fields sharded over the GPUs)")
    ap.add_argument("--fields", type=int, default=512, help="ensemble workload: fields per GPU and pass (weak scaling)")
    ap.add_argument("--field-batch", type=int, default=64, help="ensemble workload: fields per model call")
    ap.add_argument("--ens-members", type=int, default=100, help="ensemble workload: members per field")
    ap.add_argument("--graph", action="store_true",
                    help="train: step through graph.GraphedTrainStep (the step captured once in a CUDA graph, replayed per batch)")
    ap.add_argument("--strong", action="store_true", help="train: --batch is the GLOBAL batch, split over the GPUs (strong scaling)")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--members", type=int, default=15, help="ELBO ensemble size M (src/main.py:136)")
    ap.add_argument("--loss", default="afcrps", choices=["afcrps", "crps", "l1", "mse+ssim"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--device-scalars", action="store_true",
                    help="model.sync_scalars = False: elbo() returns its reconstruction terms as device tensors")
    ap.add_argument("--strict-scalars", action="store_true",
                    help="model.sync_scalars = True: real Python floats via a blocking .item() between forward and backward, "
                         "exactly as the reference (default: 'lazy' float-likes, same values, fetched on first use)")
    ap.add_argument("--no-aux", action="store_true", help="skip the roofline sweep / cpu baseline / ensemble aux")
    ap.add_argument("--cpu-batch", type=int, default=2)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/probunet_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples before this point (warm-up) are ignored."""
        try:
            self.skip = sum(1 for _ in open(self.path))
        except Exception:
            self.skip = 0

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()

    def read(self):
        """Summary of the samples taken since mark() (nvidia-smi keeps running: attaching / detaching an NVML
        client next to a timed region perturbs it)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for li, line in enumerate(open(self.path)):
            if li < getattr(self, "skip", 0):
                continue
            f = [c.strip() for c in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (restatement of the reference pinned by golden vectors)
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(args, steps, warmup, batch):
    """Times the oracle's training step (ELBO fwd + bwd + torch AdamW) on the host cores."""
    from helpers import canonical_model
    from oracle import probunet_oracle as O
    from climex_synth import make_fields
    torch.set_num_threads(os.cpu_count() or 1)
    m = canonical_model(latent_dim=args.latent)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4)
    cfg = O.ProbUNetCfg(latent_dim=args.latent)
    f = make_fields(batch, args.res, args.res, 16 if args.res >= 128 else 8, seed=1234 + 3)
    x, y = f["inputs"], f["targets"]
    g = torch.Generator().manual_seed(44)
    M = args.members if args.loss in ("afcrps", "crps") else 1
    enc, dec = O.unet_topology(cfg.unet())
    keys = [(b.key, b.cout, (b.up, b.down)) for b in enc + dec if not b.is_conv]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        eps = torch.randn(M, batch, args.latent, generator=g)
        # train-mode dropout (the reference trains with p = 0.1): Bernoulli masks drawn per block
        masks, h = {}, args.res
        for k, c, (up, down) in keys:
            h = h * 2 if up else (h // 2 if down else h)
            masks[k] = torch.rand(batch, c, h, h, generator=g) >= 0.1
        opt.zero_grad()
        out = O.elbo(full, cfg, x, y, eps, args.loss, drop_masks=masks)
        out[0].backward()
        opt.step()
        _ = float(out[0])
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.cpu_batch
    times = cpu_train_steps(args, args.steps, args.warmup, b)
    ms = 1e3 * sum(times) / len(times)
    val = b / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args), per_gpu_batch=b, workload_per_gpu_batch=args.batch,
                       sample=f"each step is one training step on a batch of {b} of the workload's {args.batch} samples (same "
                              "model, loss, resolution, optimizer); samples/s is per sample, so the figures are comparable"),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} train steps of batch {b} (same model/loss/resolution; oracle/probunet_oracle.py + torch.optim.AdamW, fp32)"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    M = args.members if args.loss in ("afcrps", "crps") else 1
    return {"workload": f"probunet_train_{args.res}x{args.res}_b{args.batch}pergpu_{args.loss}_M{M}_L{args.latent}",
            "reference_config": "BASELINE.json configs[2]: Prob U-Net training 128x128 batch 64 bf16, data-parallel",
            "per_gpu_batch": args.batch, "resolution": args.res, "elbo_members": M, "loss": args.loss,
            "latent_dim": args.latent, "optimizer": "AdamW lr 1e-4 (fused)", "dropout": 0.1,
            "elbo_scalars": ("device tensors" if getattr(args, "device_scalars", False) else
                             "python floats via blocking .item() (reference)" if getattr(args, "strict_scalars", False) else
                             "lazy float-likes (model.sync_scalars = 'lazy': async copy to pinned host memory inside elbo, awaited on first use)"),
            "l2_policy": "working set per step (saved activations, several GB) >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions)
# ------------------------------------------------------------------------------------------------
def conv_inventory(model, B, H):
    """Every conv launch of one training step: (kind, c0, c1, cout, res, ks, dtype) -> count.  dtype is the one the
    step really runs the layer in: the U-Net in bf16, the two Gaussian encoders in tf32 (f32 storage, kind::tf32)."""
    from networks import UNetBlock
    inv = {}

    def add(kind, c0, c1, cout, r, ks, dt):
        inv[(kind, c0, c1, cout, r, ks, dt)] = inv.get((kind, c0, c1, cout, r, ks, dt), 0) + 1

    def conv3(cin, cout, r, ks=3, c1=0, dgrad=True, dt="bf16"):
        add("fwd", cin - c1, c1, cout, r, ks, dt)
        add("wgrad", cin - c1, c1, cout, r, ks, dt)
        if dgrad:
            add("fwd", cout, 0, cin, r, ks, dt)      # data gradient runs the forward kernel on transposed weights

    r, c = H, model.unet.in_channels
    skips = []
    for is_dec, md in [(False, v) for v in model.unet.enc.values()] + [(True, v) for v in model.unet.dec.values()]:
        if not isinstance(md, UNetBlock):
            conv3(md.in_channels, md.out_channels, r, dgrad=False); c = md.out_channels
            skips.append(c); continue
        c1 = 0
        if is_dec and c != md.in_channels:
            c1 = skips.pop()
        if md.down: r //= 2
        if md.up: r *= 2
        conv3(md.in_channels, md.out_channels, r)
        conv3(md.out_channels, md.out_channels, r)
        if md.skip is not None and md.skip.weight is not None:
            conv3(md.in_channels, md.out_channels, r, ks=1, c1=c1)
        c = md.out_channels
        if not is_dec:
            skips.append(c)
    conv3(c, model.unet.out_channels, r)
    import _native as N
    enc = N.resolve_encoder_dtype(model.unet.compute_dtype)
    enc_dt = {N.TF32: "tf32", N.TF32_BF16S0: "tf32", N.BF16: "bf16", N.F32: "f32"}[enc]
    for encmod in (model.prior, model.posterior):
        rr, cc = H, encmod.input_channels
        for i, nf in enumerate(encmod.num_filters):
            if i: rr //= 2
            for k in range(3):
                dt = "bf16" if (enc == N.TF32_BF16S0 and i == 0 and len(encmod.num_filters) > 1) else enc_dt
                conv3(cc, nf, rr, dgrad=not (i == 0 and k == 0), dt=dt); cc = nf
    return inv


def conv_replay(model, B, H, pk):
    """Every distinct tcgen05 conv launch of the step replayed ALONE in the dtype the step runs it in, L2 flushed,
    median of 3 (isolated timings: burst peak as the denominator)."""
    import _native as N
    inv = conv_inventory(model, B, H)
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    tot_t, tot_f = 0.0, 0.0
    detail = []
    g = torch.Generator(device="cuda").manual_seed(0)
    for (kind, c0, c1, cout, r, ks, dt), cnt in sorted(inv.items()):
        if (c0 + c1) % 32 or cout % 32 or dt == "f32":
            continue                                  # first/last tiny-channel layers run on the SIMT kernels
        ndt = N.BF16 if dt == "bf16" else N.TF32
        cast = (lambda t: t.bfloat16()) if dt == "bf16" else (lambda t: t.float())
        x0 = cast(torch.randn(B, r, r, c0, device="cuda", generator=g))
        x1 = cast(torch.randn(B, r, r, c1, device="cuda", generator=g)) if c1 else None
        flops = 2.0 * B * r * r * (c0 + c1) * cout * ks * ks
        ts = []
        if kind == "fwd":
            w = cast(torch.randn(ks * ks, cout, c0 + c1, device="cuda", generator=g))
            y = torch.empty(B, r, r, cout, device="cuda", dtype=x0.dtype)
            fn = lambda: N.conv2d_nhwc(x0, w, None, x1=x1, ksize=ks, out=y, dtype=ndt)
        else:
            dy = cast(torch.randn(B, r, r, cout, device="cuda", generator=g))
            fn = lambda: N.conv2d_wgrad_nhwc(x0, dy, ks, x1=x1, want_bias=False, dtype=ndt)
        fn()
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = statistics.median(ts)
        tot_t += t * cnt; tot_f += flops * cnt
        detail.append({"kind": kind, "c0": c0, "c1": c1, "cout": cout, "res": r, "ks": ks, "dtype": dt, "count": cnt,
                       "us": round(t * 1e6, 1), "tflops": round(flops / t / 1e12, 1)})
    burst = pk["bf16_tflops"]
    ach = tot_f / tot_t / 1e12
    return {"achieved": ach, "peak": burst, "frac": ach / burst, "conv_time_per_step_ms": tot_t * 1e3,
            "conv_gflop_per_step": tot_f / 1e9,
            "how": "each distinct conv launch of the step replayed alone through the C ABI in its real dtype (U-Net bf16, "
                   "Gaussian encoders tf32) with CUDA events on the launching stream, L2 flushed (256 MiB write) before every "
                   "timed launch, median of 3; denominator = MEASURED_PEAKS.json bf16_tflops (burst: isolated timings)"}, detail


CONV_FAMILY = ("conv_halo_kernel", "conv_tc_kernel", "wgrad_tc_kernel")


def in_step_kernel_times(step_fn, N):
    """Per-kernel device time of ONE real step (CUPTI activity records through torch.profiler), with programmatic
    dependent launch switched off for that step so that kernel durations do not overlap."""
    from torch.profiler import profile, ProfilerActivity
    N.lib().pub_debug_option(b"pdl", 0)
    saved_stream_mode = os.environ.get("PROBUNET_B200_ENCODER_STREAM")
    os.environ["PROBUNET_B200_ENCODER_STREAM"] = "0"     # one stream: concurrent kernels would stretch each other's durations
    try:
        step_fn(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_fn()
            torch.cuda.synchronize()
    finally:
        N.lib().pub_debug_option(b"pdl", 1)
        if saved_stream_mode is None:
            os.environ.pop("PROBUNET_B200_ENCODER_STREAM", None)
        else:
            os.environ["PROBUNET_B200_ENCODER_STREAM"] = saved_stream_mode
    agg, first, last = {}, None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        nm = ev.name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        nm = nm.split("(")[0].replace("void ", "").replace("pub::", "")
        c = agg.setdefault(nm, [0, 0.0]); c[0] += 1; c[1] += dur
        tr = ev.time_range
        first = tr.start if first is None else min(first, tr.start)
        last = tr.end if last is None else max(last, tr.end)
    return agg, (last - first) / 1e3 if first is not None else None


def conv_roofline(model, B, H, pk, pk_kind, step_fn, N):
    """roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions): algorithmic FLOPs of all conv
    launches of one step / their device time INSIDE a real step, against the sustained cuBLAS bf16 peak."""
    replay, detail = conv_replay(model, B, H, pk)
    agg, span_ms = in_step_kernel_times(step_fn, N)
    fam = {k: v for k, v in agg.items() if any(k.startswith(f) for f in CONV_FAMILY)}
    conv_ms = sum(v[1] for v in fam.values()) / 1e3
    total_ms = sum(v[1] for v in agg.values()) / 1e3
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach = replay["conv_gflop_per_step"] / conv_ms if conv_ms > 0 else 0.0      # GFLOP / ms == TFLOP/s
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)",
            "kernel": "conv_halo_kernel / conv_tc_kernel / wgrad_tc_kernel (tcgen05 implicit GEMM; bf16 U-Net + tf32 Gaussian "
                      "encoders), all launches of one training step",
            "how": "achieved = sum over the step's tcgen05 conv launches of 2*B*H*W*Cin*Cout*k*k / their summed device time "
                   "inside ONE real training step (CUPTI kernel records via torch.profiler; programmatic dependent launch "
                   "and the encoder side streams are off for that step so kernel durations do not overlap)",
            "conv_time_in_step_ms": conv_ms, "kernel_time_in_step_ms": total_ms, "profiled_step_span_ms": span_ms,
            "share_of_step": conv_ms / total_ms if total_ms else None,
            "conv_gflop_per_step": replay["conv_gflop_per_step"], "conv_launches_in_step": sum(v[0] for v in fam.values()),
            "isolated_replay": replay,
            "step_kernels_top": [{"kernel": k[:60], "launches": v[0], "ms": round(v[1] / 1e3, 3)} for k, v in top]}, detail


def gpu_eager_baseline(args, steps=3):
    """Informational: the oracle's training step (functional torch restatement of the reference: cuDNN / cuBLAS / ATen
    kernels, the reference's own O(M^2) afCRPS) on the SAME GPU -- what `python main.py` of the reference would run at
    best on this box -- in fp32 (TF32 off) and under bf16 autocast."""
    from helpers import canonical_model
    from oracle import probunet_oracle as O
    from climex_synth import make_fields
    B, R, L = args.batch, args.res, args.latent
    M = args.members if args.loss in ("afcrps", "crps") else 1
    m = canonical_model(latent_dim=L)
    sd = {k: v.detach().clone().cuda() for k, v in m.state_dict().items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "resample_filter" not in k}
    full = dict(sd); full.update(leaves)
    opt = torch.optim.AdamW(list(leaves.values()), lr=1e-4, fused=True)
    cfg = O.ProbUNetCfg(latent_dim=L)
    f = make_fields(B, R, R, 16 if R >= 128 else 8, seed=1237)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    enc, dec = O.unet_topology(cfg.unet())
    keys = [(b.key, b.cout, (b.up, b.down)) for b in enc + dec if not b.is_conv]
    g = torch.Generator(device="cuda").manual_seed(1)
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for mode in ("fp32", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
            torch.backends.cudnn.benchmark = True

            def step():
                eps = torch.randn(M, B, L, device="cuda", generator=g)
                masks, h = {}, R
                for k, c, (up, down) in keys:
                    h = h * 2 if up else (h // 2 if down else h)
                    masks[k] = torch.rand(B, c, h, h, device="cuda", generator=g) >= 0.1
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                    o = O.elbo(full, cfg, x, y, eps, args.loss, drop_masks=masks)
                o[0].backward()
                opt.step()
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = (f"oracle/probunet_oracle.py (PyTorch eager: cuDNN/cuBLAS/ATen) training step on this GPU, batch {B}, "
                   f"{R}x{R}, {args.loss} M={M}, torch.optim.AdamW(fused); informational -- not the reference arm")
    del leaves, full, sd, opt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# ensemble workload (BASELINE.json configs[3]): M prior members per field + residual_to_hr + inverse transforms +
# CRPS / MAE per (field, variable); fields sharded over the ranks, no collective except the final gather of [T,3]
# ------------------------------------------------------------------------------------------------
ENS_FIELD_GFLOP = 28.757       # U-Net + prior forward, once per field
ENS_MEMBER_MFLOP = 36.7        # fcomb layers 1-2 per member (layer-0 feature half hoisted)
ENS_MEMBER_BYTES = 207e3       # algorithmic minimum per member: 3*H*W*4 B written + features read once


def ensemble_measure(args, model, dist, rank, world, N, fields, steps, warmup):
    from climex_synth import make_fields
    from parallel import gather_scores
    R, Mm, FB = args.res, args.ens_members, min(args.field_batch, fields)
    T = fields
    ff = make_fields(T, R, R, 16 if R >= 128 else 8, seed=99 + rank)
    sig = ff["std_hr"]
    # truth in real units (setup, not timed): residual_to_hr of the true residual + the inverse variable transforms
    hr_t = ff["lrinterp"] + ff["targets"] * (sig + 1e-10)
    sp = lambda v: torch.where(v > 20.0, v, torch.log(torch.exp(v) + 1.0) - 1e-7)   # noqa: E731
    hr_real = torch.stack([sp(hr_t[:, 0]) * 24 * 60 * 60, hr_t[:, 1] - 273.15, sp(hr_t[:, 2]) + hr_t[:, 1] - 273.15], dim=1)
    host = {k: v.contiguous().pin_memory() for k, v in (("x", ff["inputs"]), ("hr", hr_real), ("li", ff["lrinterp"]))}
    dev = {k: v.cuda() for k, v in host.items()}
    sd_ = sig.cuda()
    scores = torch.empty(T, 6, device="cuda")

    nstreams = max(1, int(os.environ.get("PROBUNET_B200_ENSEMBLE_STREAMS", "2")))

    def one_pass(src, h2d):
        # the public call: field batches alternate over two streams (host fields are copied per batch on its stream)
        crps, mae = model.sample_and_score(src["x"], Mm, src["hr"], src["li"], sd_, field_batch=FB, streams=nstreams)
        scores[:, :3], scores[:, 3:] = crps, mae
        return gather_scores(scores, [T] * world) if world > 1 else scores        # the only collective

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, warmup)):
        one_pass(dev, False)
    barrier()
    l0 = N.lib().pub_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        allv = one_pass(dev, False)
        ev[i + 1].record()
    barrier()
    launches = N.lib().pub_launch_count() - l0
    t_dev = ev[0].elapsed_time(ev[-1]) * 1e-3
    each = [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(steps)]
    one_pass(host, True).cpu()                     # warm-up of the end-to-end path
    barrier()
    w0 = time.perf_counter()
    for _ in range(steps):
        res = one_pass(host, True).cpu()           # D2H of the [T*world, 6] scores: the result the caller reads
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    members = T * Mm * world * steps
    floor_s = max((ENS_FIELD_GFLOP * 1e9 * T + ENS_MEMBER_MFLOP * 1e6 * T * Mm) / (1e12 * 1356.1),
                  ENS_MEMBER_BYTES * T * Mm / (6544e9))
    return {"members_per_s": members / t_dev, "e2e_members_per_s": members / t_e2e, "fields_per_gpu": T, "members": Mm,
            "field_batch": FB, "ms_per_pass": 1e3 * t_dev / steps, "ms_per_pass_each": each,
            "e2e_ms_per_pass": 1e3 * t_e2e / steps, "gpu_launches": int(launches),
            "h2d_bytes_per_pass": int(sum(v.numel() * 4 for v in host.values())), "d2h_bytes_per_pass": int(T * world * 6 * 4),
            "mean_crps": [round(float(v), 5) for v in allv[:, :3].mean(dim=0)],
            "roofline_floor_ms_per_pass": 1e3 * floor_s,
            "frac_of_floor": floor_s / (t_dev / steps),
            "includes": "per field: U-Net + prior once; per member: Philox rsample, fcomb, residual_to_hr + inverse transforms, "
                        "CRPS (sorted form) + MAE; per pass: one gather of the [T,6] scores"}


def run_ensemble(args, dist, rank, world, local, N, pk, pk_kind):
    from helpers import canonical_model
    model = canonical_model(latent_dim=args.latent, compute_dtype=args.dtype, device="cuda")
    model.eval()
    N.manual_seed(1000 + rank)
    sampler = ClockSampler(local); sampler.start()
    ensemble_measure(args, model, dist, rank, world, N, min(args.fields, 64), 1, 1)     # allocator / library warm-up
    sampler.mark()
    e = ensemble_measure(args, model, dist, rank, world, N, args.fields, args.steps, max(args.warmup, 3))
    clocks = sampler.read()
    sampler.stop()
    T, Mm, R = args.fields, args.ens_members, args.res
    line = {
        "metric": "ensemble_members_per_s", "value": e["members_per_s"], "unit": "members/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": e["ms_per_pass"],
        "ms_per_step_each": e["ms_per_pass_each"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"probunet_ensemble_{R}x{R}_T{T}pergpu_M{Mm}_L{args.latent}",
                   "reference_config": "BASELINE.json configs[3]: prior-sampling ensemble, 100 members per low-res field + CRPS, "
                                       "sample-parallel", "fields_per_gpu": T, "members": Mm, "field_batch": e["field_batch"],
                   "resolution": R, "latent_dim": args.latent,
                   "step": "one pass over the rank's fields: sample_and_score per field batch + one gather of the scores",
                   "l2_policy": f"members of one field batch ({e['field_batch']} x {Mm} x 3 x {R} x {R} f32) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": e["e2e_members_per_s"], "unit": "members/s", "h2d_bytes_per_step": e["h2d_bytes_per_pass"],
                "d2h_bytes_per_step": e["d2h_bytes_per_pass"], "ms_per_step": e["e2e_ms_per_pass"]},
        "gpu_launches": e["gpu_launches"], "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": (ENS_FIELD_GFLOP * T + ENS_MEMBER_MFLOP * 1e-3 * T * Mm) / e["ms_per_pass"],
                     "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "unit": "TFLOP/s", "traffic": None,
                     "kernel": "whole pass: conv_halo_kernel family (U-Net + prior once per field) + fcomb_fwd per member",
                     "how": "algorithmic FLOPs per pass (28.757 GFLOP per field + 36.7 MFLOP per member, SURVEY.md 8d) / pass time; "
                            "the per-member HBM floor (207 kB) is reported as roofline_floor"},
        "ensemble": e,
    }
    line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[1]: deterministic U-Net (src/deterministic_unet_main.py:52 -> networks.UNet defaults:
# model_channels 16, channel_mult [1,4,8,16], 14.79 M parameters, 66.25 GFLOP per training sample at 128^2) trained
# with MSE (src/trainmodel.py:158-160) and AdamW
# ------------------------------------------------------------------------------------------------
def run_detunet(args, dist, rank, world, local, N, pk):
    import networks
    from helpers import dezero
    from climex_synth import make_fields
    from optim import FusedAdamW
    from parallel import GradSynchronizer
    B, R = (args.batch if args.batch != 64 else 16), args.res          # src/trainmodel.py:36 batch 16 (override with --batch)
    torch.manual_seed(42)
    net = networks.UNet(img_resolution=(R, R), in_channels=3, out_channels=3, label_dim=0, use_diffuse=False,
                        compute_dtype=args.dtype)
    dezero(net)
    net = net.cuda().train()
    N.manual_seed(1000 + rank)
    opt = FusedAdamW(net.parameters(), lr=1e-4, grad_scale=1.0 / world)
    GradSynchronizer().install()
    f = make_fields(B, R, R, 16 if R >= 128 else 8, seed=1234 + 1 + rank)
    xh, yh = f["inputs"].pin_memory(), f["targets"].pin_memory()
    x, y = xh.cuda(), yh.cuda()

    def step(xd, yd):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(net(xd, class_labels=None), yd)   # nn.MSELoss on a [B,3,H,W] tensor: glue
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(x, y)
    barrier(); sampler.mark()
    l0 = N.lib().pub_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(x, y)
    e1.record(); barrier()
    launches = N.lib().pub_launch_count() - l0
    clocks = sampler.read()
    t_dev = e0.elapsed_time(e1) * 1e-3
    w0 = time.perf_counter()
    for _ in range(args.steps):
        lv = step(xh.cuda(non_blocking=True), yh.cuda(non_blocking=True)).item()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - w0
    sampler.stop()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    n = B * world * args.steps
    val = n / t_dev
    line = {"metric": "train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"deterministic_unet_train_{R}x{R}_b{B}pergpu_mse",
                       "reference_config": "BASELINE.json configs[1]: deterministic U-Net baseline (deterministic_unet_main.py) "
                                           "training on 128x128 pr/tasmin/tasmax", "per_gpu_batch": B, "resolution": R,
                       "model": "networks.UNet(model_channels 16, channel_mult [1,4,8,16]), 14.79 M parameters",
                       "optimizer": "AdamW lr 1e-4 (fused)", "dropout": 0.1,
                       "note": "the 16-channel 128^2 level and the 3-channel ends run on the fp32-FMA conv kernels (C % 32 != 0)"},
            "e2e": {"value": n / t_e2e, "unit": "samples/s", "h2d_bytes_per_step": int(xh.numel() * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "model_tflops_per_gpu": val / world * 66.25 * (R / 128.0) ** 2 / 1e3,
            "final_loss": float(loss.detach())}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: posterior latent exploration sweep (src/latent_exploration_posterior.py:255-343) on 256x256
# grids: per latent_dim, posterior means of N fields, unet(x0) once, two 10x10 grids decoded through fcomb on an
# expanded (non-contiguous) feature view.  PCA stays on the host (sklearn) and is not timed.
# ------------------------------------------------------------------------------------------------
def run_latent256(args, dist, rank, world, local, N, pk):
    from helpers import canonical_model
    from climex_synth import make_fields
    R = 256 if args.res == 128 else args.res
    NF, FB = 64, 16                                   # fields per latent_dim, fields per posterior call
    rows = []
    f = make_fields(NF, R, R, 16, seed=77 + rank)
    x, y = f["inputs"].cuda(), f["targets"].cuda()
    tot_t, tot_units = 0.0, 0
    for L in (2, 8, 16, 32, 64):
        m = canonical_model(latent_dim=L, compute_dtype=args.dtype, device="cuda")
        m.eval()

        def sweep():
            with torch.no_grad():
                mus = torch.cat([m.posterior(x[i:i + FB], y[i:i + FB]).base_dist.loc for i in range(0, NF, FB)])   # :255-263
                feat = m.unet(x[:1])                                                                           # :290
                grids = []
                for _ in range(2):                                                                             # deciles, +-3 sigma
                    zs = mus[:1] + torch.linspace(-3, 3, 100, device="cuda").unsqueeze(1) * mus.std(dim=0, keepdim=True)
                    grids.append(m.fcomb(feat.expand(100, -1, -1, -1), zs))                                    # :308-343
            return mus, grids
        for _ in range(2):
            sweep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            mus, grids = sweep()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        rows.append({"latent_dim": L, "ms_per_sweep": round(ms, 3), "posterior_fields_per_s": NF / ms * 1e3,
                     "decodes_per_sweep": 200})
        tot_t += ms * 1e-3; tot_units += NF
        del m
        N.workspaces.clear(); torch.cuda.empty_cache()
    line = {"metric": "posterior_fields_per_s", "value": tot_units / tot_t, "unit": "fields/s", "n_gpus": 1, "steps": args.steps,
            "warmup": 2, "ms_per_step": 1e3 * tot_t / 5, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"posterior_latent_sweep_{R}x{R}_N{NF}_latent_2_8_16_32_64",
                       "reference_config": "BASELINE.json configs[4]: posterior latent exploration sweep over latent_dim on 256x256 grids",
                       "fields": NF, "field_batch": FB, "resolution": R,
                       "step": "per latent_dim: posterior means of 64 fields + unet(x0) + 2 x 100 fcomb decodes on an expanded feature view"},
            "per_latent_dim": rows}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as G
    if rank == 0:
        G.build()
    if world > 1:
        dist.barrier()
    import _native as N
    from helpers import canonical_model
    from climex_synth import make_fields
    from optim import FusedAdamW
    from parallel import GradSynchronizer
    N.lib()
    pk, pk_kind = peaks()

    if args.workload == "ensemble":
        return run_ensemble(args, dist, rank, world, local, N, pk, pk_kind)
    if args.workload == "detunet":
        return run_detunet(args, dist, rank, world, local, N, pk)
    if args.workload == "latent256":
        return run_latent256(args, dist, rank, world, local, N, pk)
    B, R = args.batch, args.res
    if args.strong:
        assert args.batch % world == 0, "--strong: the global batch must be divisible by the number of GPUs"
        B = args.batch // world
    M = args.members if args.loss in ("afcrps", "crps") else 1
    model = canonical_model(latent_dim=args.latent, loss_type=args.loss, compute_dtype=args.dtype, device="cuda")
    model.train()                                      # dropout on, as the reference trains
    # elbo()'s reconstruction terms: "lazy" float-likes (copy to the host enqueued inside elbo, awaited on first use) by
    # default; --strict-scalars = the reference's blocking .item() between forward and backward; --device-scalars = tensors
    model.sync_scalars = False if args.device_scalars else (True if args.strict_scalars else "lazy")
    N.manual_seed(1000 + rank)
    opt = FusedAdamW(model.parameters(), lr=1e-4, grad_scale=1.0 / world)
    sync = GradSynchronizer().install()
    f = make_fields(B, R, R, 16 if R >= 128 else 8, seed=1234 + 3 + rank)
    x_host, y_host = f["inputs"].pin_memory(), f["targets"].pin_memory()
    x, y = x_host.cuda(), y_host.cuda()

    def step_eager(xd, yd):
        opt.zero_grad(set_to_none=True)
        out = model.elbo(xd, yd, None, M=M) if args.loss in ("afcrps", "crps") else model.elbo(xd, yd, None)
        out[0].backward()
        opt.step()
        return out[0]

    step_device = step_eager
    if args.graph:
        from graph import GraphedTrainStep
        gmain = GraphedTrainStep(model, opt, x, y, M=M if args.loss in ("afcrps", "crps") else None, warmup=3)
        step_device = lambda xd, yd: gmain(xd, yd)[0]       # noqa: E731  (copies the batch into the graph's static inputs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started BEFORE the warm-up: its NVML start-up (~1 s) briefly serialises with the CUDA driver
    # and must not fall into the timed region
    sampler = ClockSampler(local); sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device(x, y)
    # ---- device-resident timing
    barrier()
    sampler.mark()
    l0 = N.lib().pub_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    c0 = time.perf_counter()
    for i in range(args.steps):
        loss = step_device(x, y)
        evs[i + 1].record()
    cpu_enqueue = (time.perf_counter() - c0) / args.steps
    barrier()
    launches = (N.lib().pub_launch_count() - l0)
    if args.graph:
        launches = gmain.launches_per_step * args.steps      # replays do not pass through the library's launch counter
    clocks = sampler.read()
    t_dev = evs[0].elapsed_time(evs[-1]) * 1e-3
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    # ---- end to end through the public API: pinned host inputs -> H2D -> step -> D2H of the loss, every step.
    # The input pipeline is the usual double buffer of a pinned-memory loader: the H2D copy of step i + 1 is enqueued on
    # a copy stream before step i, so it runs on the copy engine while step i computes; the loss of every step is copied
    # to pinned host memory when the step is enqueued and read by the host one step later (logging with a one-step lag:
    # the host never drains the GPU between steps).  All n losses are read inside the timed region.
    copy_stream = torch.cuda.Stream()
    slots = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])          # the step that last read this slot has been enqueued and ran
            slots[k][0].copy_(x_host, non_blocking=True)
            slots[k][1].copy_(y_host, non_blocking=True)
            ready[k].record(copy_stream)

    def e2e_loop(n):
        from lazy_scalar import LazyScalar
        each, losses, prev = [], [], None
        main = torch.cuda.current_stream()
        for k in range(2):
            consumed[k].record(main)
        prefetch(0)
        for i in range(n):
            w1 = time.perf_counter()
            if i + 1 < n:
                prefetch(i + 1)
            main.wait_event(ready[i & 1])
            lossv = step_device(*slots[i & 1])
            consumed[i & 1].record(main)
            cur = LazyScalar(lossv)                 # D2H of this step's loss into pinned memory, enqueued now ...
            if prev is not None:
                losses.append(float(prev))          # ... and read once the NEXT step has been enqueued (one-step lag)
            prev = cur
            each.append(round(1e3 * (time.perf_counter() - w1), 2))
        losses.append(float(prev))
        assert len(losses) == n and all(v == v and abs(v) != float("inf") for v in losses)
        return each

    e2e_loop(2)             # warm-up of THIS path (first pinned H2D + allocator growth cost ~60 ms once)
    barrier()
    w0 = time.perf_counter()
    e2e_each = e2e_loop(args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - w0
    sampler.stop()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    samples = B * world * args.steps
    value, e2e = samples / t_dev, samples / t_e2e
    gflop = GFLOP_TRAIN_SAMPLE_M1 * (R / 128.0) ** 2 + GFLOP_PER_EXTRA_MEMBER * (M - 1) * (R / 128.0) ** 2

    line = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps,
        "ms_per_step_each": [round(v, 2) for v in per_step], "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": dict(workload_config(args), per_gpu_batch=B, global_batch=B * world,
                       stepping="CUDA graph replay (graph.GraphedTrainStep)" if args.graph else "eager (model.elbo + backward + FusedAdamW.step)"),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * t_e2e / args.steps, "ms_per_step_each": e2e_each,
                "input_pipeline": "double-buffered pinned H2D on a copy stream (step i + 1's copy overlaps step i); every step's loss is copied to pinned host memory and read by the host one step later"},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": 1e3 * cpu_enqueue,
        "clocks": clocks,
        "model_tflops_per_gpu": value / world * gflop / 1e3,
        "final_loss": float(loss.detach()),
        "grad_sync": {"collectives_per_step": sync.calls // max(1, (max(args.warmup, 3) + 2 * args.steps)),
                      "bytes_per_step": sync.bytes // max(1, (max(args.warmup, 3) + 2 * args.steps)), "world": world},
    }
    if args.graph:
        args.no_aux = True          # the auxiliary legs re-use the eager step; run them without --graph
    # The timed regions are over.  The auxiliary legs below run training steps on rank 0 ONLY (in-step roofline, CUDA
    # graph): with the synchronizer still installed those steps would issue NCCL all-reduces no other rank joins.
    sync.wait_all()
    GradSynchronizer.uninstall()
    if rank == 0 and not args.no_aux:
        # ---- roofline of the dominant kernel family, measured live
        try:
            roof, detail = conv_roofline(model, B, R, pk, pk_kind, lambda: step_device(x, y), N)
            # DRAM traffic of the same kernel family over one step: from the ncu launch list of THIS round's build
            # (profiles/r02_launches_step.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch, summed); ncu cannot
            # run inside the bench, so the value is only reported when that file exists and says which commit it is from
            try:
                fam = json.load(open(os.path.join(ROOT, "profiles", "r02_launches_step_family.json")))
                cf = fam["conv_family"]
                roof["traffic"] = cf["dram_bytes"]
                roof["traffic_source"] = (f"profiles/r02_launches_step.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                          f"{cf['launches']} conv-family launches of one step, build {fam.get('build', '?')}; "
                                          f"ncu share of step {cf['share_of_step']:.3f})")
            except Exception:
                roof["traffic"] = None
            line["roofline"] = roof
            line["roofline_detail_top"] = sorted(detail, key=lambda d: -d["us"] * d["count"])[:8]
        except Exception as ex:  # never lose the headline line
            line["roofline"] = {"error": repr(ex)}
    if not args.no_aux and world == 1:
        # ---- the same step captured once in a CUDA graph and replayed (graph.GraphedTrainStep; single GPU only here:
        # the multi-GPU capture with NCCL nodes is validated at 16 samples per GPU)
        try:
            from graph import GraphedTrainStep
            loss = None               # a live loss keeps the eager steps' autograd graph (bound to the default stream) alive
            gstep = GraphedTrainStep(model, opt, x, y, M=M if args.loss in ("afcrps", "crps") else None, warmup=2)
            for _ in range(3):
                gstep(x, y)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            h0 = time.perf_counter()
            for _ in range(args.steps):
                gout = gstep(x, y)
            host_ms = 1e3 * (time.perf_counter() - h0) / args.steps
            g1.record()
            torch.cuda.synchronize()
            tg = g0.elapsed_time(g1) * 1e-3
            line["cuda_graph"] = {"samples_per_s": B * args.steps / tg, "ms_per_step": 1e3 * tg / args.steps,
                                  "host_enqueue_ms_per_step": host_ms, "launches_per_step": gstep.launches_per_step,
                                  "what": "graph.GraphedTrainStep: elbo + backward + fused AdamW captured once (encoder side "
                                          "streams included), replayed per step; device-side step counter and random salt; "
                                          "ELBO scalars stay on the device"}
            gstep.close()
            del gstep, gout
        except Exception as ex:
            line["cuda_graph"] = {"error": repr(ex)[:300]}
    if not args.no_aux:
        # ---- second half of BASELINE.json's metric: ensemble members/s (configs[3]), every rank its own fields, one
        # final gather -- a short run here; `--workload ensemble` is the full-length line
        try:
            model.eval()
            ens = ensemble_measure(args, model, dist, rank, world, N, fields=min(args.fields, 128), steps=3, warmup=1)
            if rank == 0:
                line["ensemble"] = ens
        except Exception as ex:
            if rank == 0:
                line["ensemble"] = {"error": repr(ex)}
    if rank == 0 and not args.no_aux and world == 1:     # the remaining extras (and the CPU baseline) at N = 1 only
        # ---- auxiliary: GPU-side dataset transform (__getitem__ math, src/climex_utils.py:197-225)
        try:
            from climex_gpu import ClimexBatchTransform
            hr_all = torch.randn(4 * B, 3, R, R, device="cuda") * 4 + 280
            tr = ClimexBatchTransform(lowres_scale=16 if R >= 128 else 8)
            tr.compute_stats(hr_all)
            for _ in range(3):
                tr(hr_all[:B])
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for i in range(20):
                tr(hr_all[(i % 4) * B:(i % 4 + 1) * B])
            f1.record(); torch.cuda.synchronize()
            dt_f = f0.elapsed_time(f1) * 1e-3 / 20
            line["feeder"] = {"samples_per_s": B / dt_f, "gb_per_s": B * 3 * R * R * 4 * 4 / dt_f / 1e9,
                              "what": "climex_transform_kernel: hr [B,3,H,W] in HBM -> inputs, targets, lrinterp, lr "
                                      "(reads 1x, writes 3x; the reference does this per sample on the host)"}
        except Exception as ex:
            line["feeder"] = {"error": repr(ex)}
        # ---- same-box PyTorch-eager baseline (informational)
        try:
            del model, opt
            torch.cuda.empty_cache()
            N.workspaces.clear()
            line["gpu_eager_baseline"] = gpu_eager_baseline(args)
        except Exception as ex:
            line["gpu_eager_baseline"] = {"error": repr(ex)}
        # ---- CPU baseline: the oracle on the host cores, bounded sample
        try:
            torch.cuda.synchronize()
            times = cpu_train_steps(args, 2, 1, args.cpu_batch)
            v = args.cpu_batch / (sum(times) / len(times))
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"2 train steps of batch {args.cpu_batch} after 1 warm-up, same model/loss/resolution, fp32 "
                                              "(oracle/probunet_oracle.py + torch.optim.AdamW)"}
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

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
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--members", type=int, default=15, help="ELBO ensemble size M (src/main.py:136)")
    ap.add_argument("--loss", default="afcrps", choices=["afcrps", "crps", "l1", "mse+ssim"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--device-scalars", action="store_true",
                    help="model.sync_scalars = False: elbo() returns its reconstruction terms as device tensors instead of "
                         "Python floats (no host sync between forward and backward); default keeps the reference's floats")
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
        "config": workload_config(args),
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
            "elbo_scalars": "device tensors" if getattr(args, "device_scalars", False) else "python floats (reference)",
            "l2_policy": "working set per step (saved activations, several GB) >> 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# roofline of the dominant kernel family (tcgen05 implicit-GEMM convolutions)
# ------------------------------------------------------------------------------------------------
def conv_inventory(model, B, H):
    """Every conv launch of one training step: (kind, c0, c1, cout, res, ks) -> count."""
    from networks import UNetBlock
    inv = {}

    def add(kind, c0, c1, cout, r, ks):
        inv[(kind, c0, c1, cout, r, ks)] = inv.get((kind, c0, c1, cout, r, ks), 0) + 1

    def conv3(cin, cout, r, ks=3, c1=0, dgrad=True):
        add("fwd", cin - c1, c1, cout, r, ks)
        add("wgrad", cin - c1, c1, cout, r, ks)
        if dgrad:
            add("fwd", cout, 0, cin, r, ks)          # data gradient runs the forward kernel on transposed weights

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
    for encmod in (model.prior, model.posterior):
        rr, cc = H, encmod.input_channels
        for i, nf in enumerate(encmod.num_filters):
            if i: rr //= 2
            for k in range(3):
                conv3(cc, nf, rr, dgrad=not (i == 0 and k == 0)); cc = nf
    return inv


def conv_roofline(model, B, H, pk, pk_kind):
    import _native as N
    inv = conv_inventory(model, B, H)
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    tot_t, tot_f, tc_t, tc_f = 0.0, 0.0, 0.0, 0.0
    detail = []
    g = torch.Generator(device="cuda").manual_seed(0)
    for (kind, c0, c1, cout, r, ks), cnt in sorted(inv.items()):
        if (c0 + c1) % 32 or cout % 32:
            continue                                  # first/last tiny-channel layers run on the SIMT kernel
        x0 = torch.randn(B, r, r, c0, device="cuda", generator=g).bfloat16()
        x1 = torch.randn(B, r, r, c1, device="cuda", generator=g).bfloat16() if c1 else None
        flops = 2.0 * B * r * r * (c0 + c1) * cout * ks * ks
        ts = []
        if kind == "fwd":
            w = torch.randn(ks * ks, cout, c0 + c1, device="cuda", generator=g).bfloat16()
            y = torch.empty(B, r, r, cout, device="cuda", dtype=torch.bfloat16)
            fn = lambda: N.conv2d_nhwc(x0, w, None, x1=x1, ksize=ks, out=y)
        else:
            dy = torch.randn(B, r, r, cout, device="cuda", generator=g).bfloat16()
            fn = lambda: N.conv2d_wgrad_nhwc(x0, dy, ks, x1=x1, want_bias=False)
        fn()
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = statistics.median(ts)
        tot_t += t * cnt; tot_f += flops * cnt
        detail.append({"kind": kind, "c0": c0, "c1": c1, "cout": cout, "res": r, "ks": ks, "count": cnt,
                       "us": round(t * 1e6, 1), "tflops": round(flops / t / 1e12, 1)})
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach = tot_f / tot_t / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)",
            "kernel": "conv_tc_kernel / conv_halo_kernel / wgrad_tc_kernel (tcgen05 implicit GEMM), all conv launches of one step",
            "how": "each distinct conv launch of the step replayed alone through the C ABI with CUDA events on the "
                   "launching stream, L2 flushed (256 MiB write) before every timed launch, median of 3; "
                   "achieved = sum(count*2*B*H*W*Cin*Cout*k*k) / sum(count*time)",
            "conv_time_per_step_ms": tot_t * 1e3, "conv_gflop_per_step": tot_f / 1e9}, detail


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

    B, R = args.batch, args.res
    M = args.members if args.loss in ("afcrps", "crps") else 1
    model = canonical_model(latent_dim=args.latent, loss_type=args.loss, compute_dtype=args.dtype, device="cuda")
    model.train()                                      # dropout on, as the reference trains
    model.sync_scalars = not args.device_scalars
    N.manual_seed(1000 + rank)
    opt = FusedAdamW(model.parameters(), lr=1e-4, grad_scale=1.0 / world)
    sync = GradSynchronizer().install()
    f = make_fields(B, R, R, 16 if R >= 128 else 8, seed=1234 + 3 + rank)
    x_host, y_host = f["inputs"].pin_memory(), f["targets"].pin_memory()
    x, y = x_host.cuda(), y_host.cuda()

    def step_device(xd, yd):
        opt.zero_grad(set_to_none=True)
        out = model.elbo(xd, yd, None, M=M) if args.loss in ("afcrps", "crps") else model.elbo(xd, yd, None)
        out[0].backward()
        opt.step()
        return out[0]

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
    clocks = sampler.read()
    t_dev = evs[0].elapsed_time(evs[-1]) * 1e-3
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    # ---- end to end through the public API: pinned host inputs -> H2D -> step -> D2H of the loss
    for _ in range(2):      # warm-up of THIS path (first pinned H2D + allocator growth cost ~60 ms once)
        xd, yd = x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True)
        step_device(xd, yd).item()
    barrier()
    w0 = time.perf_counter()
    e2e_each = []
    for _ in range(args.steps):
        w1 = time.perf_counter()
        xd, yd = x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True)
        lv = step_device(xd, yd).item()
        e2e_each.append(round(1e3 * (time.perf_counter() - w1), 2))
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
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * t_e2e / args.steps, "ms_per_step_each": e2e_each},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": 1e3 * cpu_enqueue,
        "clocks": clocks,
        "model_tflops_per_gpu": value / world * gflop / 1e3,
        "final_loss": float(loss.detach()),
        "grad_sync": {"collectives_per_step": sync.calls // max(1, (max(args.warmup, 3) + 2 * args.steps)),
                      "bytes_per_step": sync.bytes // max(1, (max(args.warmup, 3) + 2 * args.steps)), "world": world},
    }
    if rank == 0 and not args.no_aux:
        # ---- roofline of the dominant kernel family, measured live
        try:
            model.eval()
            roof, detail = conv_roofline(model, B, R, pk, pk_kind)
            roof["share_of_step"] = roof["conv_time_per_step_ms"] / line["ms_per_step"]
            # DRAM traffic of the same kernel family over one step, from the committed ncu launch list
            # (profiles/r01d_launches_step.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch, summed)
            try:
                fam = json.load(open(os.path.join(ROOT, "profiles", "r01d_launches_step_family.json")))["conv_family"]
                roof["traffic"] = fam["dram_bytes"]
                roof["traffic_source"] = ("profiles/r01d_launches_step.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                          f"{fam['launches']} conv-family launches of one step; ncu share of step {fam['share_of_step']:.3f})")
            except Exception:
                pass
            line["roofline"] = roof
            line["roofline_detail_top"] = sorted(detail, key=lambda d: -d["us"] * d["count"])[:8]
        except Exception as ex:  # never lose the headline line
            line["roofline"] = {"error": repr(ex)}
    if rank == 0 and not args.no_aux and world == 1:     # the remaining extras (and the CPU baseline) at N = 1 only
        # ---- auxiliary: ensemble members/s (BASELINE config 4 shape: M=100 prior members per field + CRPS/MAE)
        try:
            import metrics as MET
            model.eval()
            T, Mm = 64, 100
            ff = make_fields(T, R, R, 16 if R >= 128 else 8, seed=99)
            xi, hr, li, sd_ = ff["inputs"].cuda(), ff["hr"].cuda(), ff["lrinterp"].cuda(), ff["std_hr"].cuda()
            for _ in range(2):
                ens = model.sample(xi, Mm)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            for i in range(5):
                ens = model.sample(xi, Mm)
                crps, mae = MET.ensemble_scores_from_residuals(ens, hr, li, sd_)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ens_ms = [round(ev[i].elapsed_time(ev[i + 1]), 2) for i in range(5)]
            line["ensemble"] = {"members_per_s": T * Mm / (statistics.median(ens_ms) * 1e-3), "fields": T, "members": Mm,
                                "ms_per_pass_each": ens_ms,
                                "includes": "unet+prior once per field, fcomb x M, residual_to_hr+CRPS+MAE kernel"}
        except Exception as ex:
            line["ensemble"] = {"error": repr(ex)}
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

"""Per-kernel time of ONE training step from CUPTI activity records (torch.profiler), aggregated by kernel
name and grid.  Cheap (seconds) next to an `ncu` replay of 1300 launches; kernels overlap as in the real step.

    python tools/step_profile.py [--batch 64] [--res 128] [--members 15] [--out gpurun_out/step_profile.txt]
"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "prob-unet-climate-downscaling_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def short(name):
    s = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    s = re.sub(r"\(.*", "", s)
    s = re.sub(r"^void ", "", s)
    return s.replace("pub::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--members", type=int, default=15)
    ap.add_argument("--loss", default="afcrps")
    ap.add_argument("--ensemble", type=int, default=0, help="profile one sample_and_score pass with this many members instead of a training step")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.txt"))
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    import _native as N
    from helpers import canonical_model
    from climex_synth import make_fields
    from optim import FusedAdamW
    torch.cuda.set_device(0)
    model = canonical_model(latent_dim=32, loss_type=args.loss, compute_dtype="bf16", device="cuda")
    model.train()
    N.manual_seed(1000)
    opt = FusedAdamW(model.parameters(), lr=1e-4)
    f = make_fields(args.batch, args.res, args.res, 16 if args.res >= 128 else 8, seed=1237)
    x, y = f["inputs"].cuda(), f["targets"].cuda()

    def step():
        opt.zero_grad(set_to_none=True)
        out = model.elbo(x, y, None, M=args.members) if args.loss != "l1" else model.elbo(x, y, None)
        out[0].backward()
        opt.step()

    if args.ensemble:
        model.eval()
        hr, li, sd_ = f["hr"].cuda(), f["lrinterp"].cuda(), f["std_hr"].cuda()

        def step():                                          # noqa: F811
            model.sample_and_score(x, args.ensemble, hr, li, sd_)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    fine = collections.defaultdict(lambda: [0, 0.0])
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        nm = short(ev.name)
        agg[nm][0] += 1; agg[nm][1] += dur
        tr = ev.time_range
        t_first = tr.start if t_first is None else min(t_first, tr.start)
        t_last = tr.end if t_last is None else max(t_last, tr.end)
    tot = sum(v[1] for v in agg.values())
    lines = [f"step span {(t_last - t_first) / 1e3:.3f} ms; sum of kernel time {tot / 1e3:.3f} ms; launches {sum(v[0] for v in agg.values())}"]
    for nm, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f"{us / 1e3:9.3f} ms  {100 * us / tot:5.1f}%  x{c:<5d} avg {us / c:8.1f} us  {nm[:100]}")
    txt = "\n".join(lines)
    print(txt)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(txt + "\n")
    try:
        prof.export_chrome_trace(args.out.replace(".txt", ".trace.json"))
    except Exception as e:  # noqa: BLE001
        print("trace export failed:", e)


if __name__ == "__main__":
    main()

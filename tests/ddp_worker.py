"""torchrun worker of tests/test_gpu_multi.py: SURVEY.md section 4 item 5 -- an N-GPU data-parallel step must equal
the 1-GPU step on the concatenated batch.  Every rank owns a different slice of the batch (and different eps), takes
`steps` AdamW steps through GradSynchronizer + FusedAdamW; checks (a) after step 1 the all-reduced gradients / world
equal the single-process gradients of the full batch, (b) parameters stay BIT-identical across ranks, (c) after the
last step they match the single-process run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "prob-unet-climate-downscaling_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    dtype = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    steps = 2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from helpers import canonical_model, rel_err
    from climex_synth import make_fields
    from optim import FusedAdamW
    from parallel import GradSynchronizer
    per, M, L = 2, 3, 32
    f = make_fields(per * world, 64, 64, 8, seed=31)
    X, Y = f["inputs"].cuda(), f["targets"].cuda()
    EPS = torch.randn(steps, M, per * world, L, generator=torch.Generator().manual_seed(32)).cuda()
    sl = slice(rank * per, (rank + 1) * per)

    def run(model, opt, x, y, eps_of_step, nsteps, grads_after_first=None):
        for s in range(nsteps):
            opt.zero_grad(set_to_none=True)
            total, _, _ = model.elbo(x, y, None, M=M, eps=eps_of_step(s))
            total.backward()
            if s == 0 and grads_after_first is not None:
                grads_after_first()
            opt.step()

    # ---- data-parallel run
    m = canonical_model(compute_dtype=dtype, device="cuda")        # eval(): dropout off, every rank the same weights
    sync = GradSynchronizer().install()
    opt = FusedAdamW(m.parameters(), lr=1e-3, grad_scale=1.0 / world)
    g_dp = {}

    def grab():
        sync.wait_all()                                            # reduced sums are now in p.grad (checked aliasing)
        for n, p in m.named_parameters():
            g_dp[n] = p.grad.detach().clone() / world
    run(m, opt, X[sl], Y[sl], lambda s: EPS[s][:, sl].contiguous(), steps, grab)
    GradSynchronizer.uninstall()
    torch.cuda.synchronize()
    # (b) bit-identical parameters on every rank
    for n, p in m.named_parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, p.detach()), f"rank {rank}: parameter {n} diverged from rank 0"
    # ---- single-process run on the concatenated batch (every rank does it; cheap at this size)
    m1 = canonical_model(compute_dtype=dtype, device="cuda")
    opt1 = FusedAdamW(m1.parameters(), lr=1e-3)
    g_1 = {}

    def grab1():
        for n, p in m1.named_parameters():
            g_1[n] = p.grad.detach().clone()
    run(m1, opt1, X, Y, lambda s: EPS[s].contiguous(), steps, grab1)
    gtol = 2e-4 if dtype == "fp32" else 3e-2
    bad = [(n, rel_err(g_dp[n], g_1[n])) for n in g_1 if float(g_1[n].norm()) > 1e-7 and rel_err(g_dp[n], g_1[n]) > gtol]
    assert not bad, f"rank {rank}: all-reduced gradients differ from the full-batch gradients: {bad[:6]}"
    ptol = 5e-3 if dtype == "fp32" else 5e-2       # Adam divides by sqrt(v): ~1e-8 gradients amplify rounding noise
    worst = max(rel_err(p, dict(m1.named_parameters())[n]) for n, p in m.named_parameters())
    assert worst < ptol, f"rank {rank}: parameters after {steps} steps differ from the single-process run ({worst})"
    dist.barrier()
    if rank == 0:
        print(f"ddp_worker ok: world={world} dtype={dtype} worst_param_rel_err={worst:.2e} "
              f"worst_grad_rel_err={max(rel_err(g_dp[n], g_1[n]) for n in g_1 if float(g_1[n].norm()) > 1e-7):.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

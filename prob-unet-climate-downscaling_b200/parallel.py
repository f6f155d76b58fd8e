"""One-process-per-GPU plumbing (torch.distributed; NCCL over NVLink on the box, gloo in CPU tests).

The reference is single-process.  The hot path shards two ways:

* training -- pure data parallel: every rank runs the full model on its slice of the batch and
  gradients are SUM-all-reduced (``GradSynchronizer``); the 1/world factor is folded into
  FusedAdamW's ``grad_scale``.  Each engine (fcomb, U-Net, posterior, prior) hands its flat
  gradient buffer to the synchronizer as soon as its backward kernels are enqueued; the collective
  runs on NCCL's stream while the NEXT sub-network's backward kernels run on the compute stream --
  the compute stream only waits for the collectives in ``wait_all()``, which FusedAdamW.step()
  (or the user, before any other use of the gradients) calls.
* ensemble sampling -- fields are partitioned over ranks, no collective until the final gather
  of the [T,3] score arrays (``shard_range`` / ``gather_scores``).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) slice of ``n_items`` for ``rank``."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradSynchronizer:
    """Sums every flat gradient buffer the engines report across ranks."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0
        self.bytes = 0
        self.pending = []          # (work handle, buffer kept alive until the wait)

    def __call__(self, flat):
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        if self.world > 1:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((work, flat))

    def wait_all(self):
        """Makes the current stream wait for every collective issued since the last call (stream-ordered for
        NCCL: no host block; blocking for gloo).  Must run before the gradients are consumed."""
        for work, _ in self.pending:
            work.wait()
        self.pending.clear()

    def install(self):
        import _native
        _native.grad_ready_callback = self
        return self

    @staticmethod
    def uninstall():
        import _native
        _native.grad_ready_callback = None


def gather_scores(local, counts, group=None):
    """All ranks' [T_local, C] score rows -> [T, C] on every rank (the only collective of the
    sample-parallel ensemble path)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    tmax = max(counts)
    pad = torch.zeros(tmax, local.shape[1], device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)

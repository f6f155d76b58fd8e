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
    """Sums every flat gradient buffer the engines report across ranks.

    Contract (checked, not assumed): the engines hand autograd VIEWS of the flat buffer as the parameter gradients; with
    ``zero_grad(set_to_none=True)`` autograd stores those views as ``p.grad`` itself, so the in-place all-reduce of the
    flat buffer IS the reduction of ``p.grad``.  ``wait_all()`` verifies that aliasing for every parameter and raises
    if autograd accumulated into an older ``p.grad`` instead (``zero_grad(set_to_none=False)`` or gradient accumulation:
    the reduced values would never reach ``p.grad`` and the ranks would silently diverge).  A new backward while the
    previous step's collectives were never consumed through ``wait_all()`` (an optimizer that does not know about the
    synchronizer) raises too.  ``FusedAdamW.step()`` calls ``wait_all()`` and folds 1/world into its update; for any
    other optimizer call ``sync.wait_all(average=True)`` before ``optimizer.step()``."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0
        self.bytes = 0
        self.pending = []          # (work handle, flat buffer kept alive until the wait, params, views)

    def __call__(self, flat, params=None, views=None):
        """Called by every engine backward with its flat gradient buffer and the parameters whose gradients are the
        consecutive slices of it.  (`views` are NOT retained: autograd only adopts an incoming gradient tensor as
        ``p.grad`` when nobody else holds a reference to it -- otherwise it clones.)"""
        if params:
            for _, _, old_params, _ in self.pending:
                if old_params and old_params[0] is params[0]:      # the same sub-network reported twice
                    self.pending.clear()
                    raise RuntimeError("GradSynchronizer: a new backward started but the previous step's gradient "
                                       "all-reduces were never consumed -- call sync.wait_all(average=True) before "
                                       "optimizer.step() (FusedAdamW does it itself)")
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        work = None
        if self.world > 1:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        had_grad = [p.grad is not None for p in params] if params else None    # before AccumulateGrad runs
        self.pending.append((work, flat, list(params) if params else None, had_grad))

    def wait_all(self, average=False):
        """Makes the current stream wait for every collective issued since the last call (stream-ordered for
        NCCL: no host block; blocking for gloo) and makes the reduced buffer authoritative: a ``p.grad`` that aliases
        its slice of the flat buffer already holds the sum; one that autograd cloned (it had been None) is overwritten
        with the reduced slice; one that existed before the backward (gradient accumulation) is an error.
        ``average=True`` (for optimizers other than FusedAdamW) divides the sums by the world size."""
        pending, self.pending = self.pending, []
        for work, flat, params, had_grad in pending:
            if work is not None:
                work.wait()
            if average and self.world > 1:
                flat.div_(self.world)
            if not params:
                continue
            off = 0
            for p, had in zip(params, had_grad):
                n = p.numel()
                if p.grad is not None and p.grad.data_ptr() == flat.data_ptr() + 4 * off:
                    off += n
                    continue                       # autograd adopted the view: reduced in place
                if had or p.grad is None:
                    raise RuntimeError(
                        "GradSynchronizer: a parameter's .grad is not the all-reduced buffer (autograd accumulated "
                        "into an existing .grad).  Use optimizer.zero_grad(set_to_none=True) every step; gradient "
                        "accumulation across backward passes is not supported with the synchronizer installed.")
                with torch.no_grad():
                    p.grad.copy_(flat[off:off + n].view(p.shape))    # autograd cloned the local values: replace them
                off += n

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

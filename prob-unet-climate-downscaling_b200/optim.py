"""Fused multi-tensor AdamW: one kernel launch for all 391 tensors.

Same update rule and defaults as ``torch.optim.AdamW`` which the reference constructs at
src/main.py:103 and steps at src/train_prob_unet_model.py:139-141 (decoupled weight decay,
bias correction, eps outside the sqrt).  ``grad_scale`` folds the 1/world_size of a
data-parallel SUM all-reduce into the update.
"""
import ctypes as C

import torch

import _native as N


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale
        self._tables = {}
        self._events = {}
        self._step_dev = None      # device int32 step counter (CUDA-graph mode, see graph.GraphedTrainStep)

    def use_device_step(self, counter):
        """counter: 1-element int32 CUDA tensor holding the optimizer step count, advanced on the device
        (pub_advance_counters) before every step -- required under CUDA-graph capture, where the host-side step count
        would be frozen into the captured kernel arguments.  None switches back to the host count."""
        if counter is not None:
            assert counter.is_cuda and counter.dtype == torch.int32 and counter.numel() >= 1
            host_steps = {g.get("step", 0) for g in self.param_groups}
            assert len(host_steps) == 1, "parameter groups with different step counts cannot share one device counter"
        self._step_dev = counter

    def _table(self, gi, plist):
        """Device table of {p, g, m, v, n} rows; refreshed only when a pointer changes.  The pinned staging buffer and
        the device buffer are allocated ONCE per group: a refresh inside a CUDA-graph capture (the gradients live at
        new addresses in the graph's memory pool) must not allocate pinned memory -- that invalidates the capture --
        so it only rewrites the staging buffer on the host and enqueues the H2D copy (a memcpy node of the graph)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plist)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent
        nbytes = C.sizeof(N.AdamWEntry) * len(plist)
        if ent is None or ent[2].numel() != nbytes:
            host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            dev = torch.empty(nbytes, dtype=torch.uint8, device=plist[0].device)
        else:
            dev, host = ent[1], ent[2]
        capturing = torch.cuda.is_current_stream_capturing()
        ev = self._events.get(gi)
        if ev is not None and not capturing:
            ev.synchronize()          # the previous H2D copy of this staging buffer must have finished before it is rewritten
        rows = (N.AdamWEntry * len(plist)).from_address(host.data_ptr())
        for i, p in enumerate(plist):
            st = self.state[p]
            rows[i].p, rows[i].g = p.data_ptr(), p.grad.data_ptr()
            rows[i].m, rows[i].v, rows[i].n = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
        dev.copy_(host, non_blocking=True)
        if not capturing:
            ev = self._events.setdefault(gi, torch.cuda.Event())
            ev.record()
        ent = (key, dev, host, max(p.numel() for p in plist))
        self._tables[gi] = ent
        return ent

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        sync = N.grad_ready_callback
        if sync is not None and hasattr(sync, "wait_all"):
            sync.wait_all()           # gradient all-reduces overlapped with the rest of backward end here
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                N.require_cuda(p)
                if not p.grad.is_contiguous() or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise N.NativeError("FusedAdamW needs contiguous f32 parameters and gradients")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            group["step"] = group.get("step", 0) + 1
            _, dev, _, maxn = self._table(gi, plist)
            b1, b2 = group["betas"]
            if self._step_dev is not None:
                N.check(N.lib().pub_adamw_step_dev(N.ptr(dev), len(plist), C.c_int64(maxn), C.c_float(group["lr"]),
                                                   C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]),
                                                   C.c_float(group["weight_decay"]), N.ptr(self._step_dev),
                                                   C.c_float(self.grad_scale), N.stream()), "pub_adamw_step_dev")
                continue
            N.check(N.lib().pub_adamw_step(N.ptr(dev), len(plist), C.c_int64(maxn), C.c_float(group["lr"]), C.c_float(b1),
                                           C.c_float(b2), C.c_float(group["eps"]), C.c_float(group["weight_decay"]),
                                           int(group["step"]), C.c_float(self.grad_scale), N.stream()), "pub_adamw_step")
        return loss

"""CUDA-graph'd training step (replaces the per-batch body of train_probunet_step,
src/train_prob_unet_model.py:133-141: elbo -> zero_grad -> backward -> optimizer.step).

The step is ~760 kernel launches issued from C++ behind ~12 autograd nodes; at the benchmark batch (64 per GPU) the
GPU is the bottleneck, but at small per-GPU batches (strong scaling: 8 samples per GPU on 8 GPUs) the ~7 ms of host
enqueue per step is.  Capturing the whole step once and replaying it takes the host out of the loop.

What a replay cannot get from captured kernel ARGUMENTS comes from device memory:
  * the optimizer step count (Adam bias corrections)  -> FusedAdamW.use_device_step + pub_adamw_step_dev
  * per-step randomness (dropout masks, rsample eps)   -> a device salt mixed into every dropout key / Philox seed
both advanced by one tiny kernel at the top of the captured step (pub_advance_counters).
The ELBO's reconstruction terms are returned as device tensors (model.sync_scalars = False): a .item() inside the step
would be a host sync and cannot be captured; read them every N steps instead (de-synced logging).
"""
import ctypes as C
import os
import sys
import time

import torch

import _native as N


def _dbg(msg):
    if os.environ.get("PROBUNET_B200_GRAPH_DEBUG"):
        print(f"[graph {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, x_example, y_example, M=15); then per batch
    ``out = step(x, y)`` -> the elbo() tuple as device tensors of the step that was just ENQUEUED (no host sync).

    ``model.beta_*`` / learning rates are frozen at capture time: call ``recapture()`` after changing them (the
    reference changes the betas once per epoch, src/main.py:122-123).  Works with a GradSynchronizer installed
    (the NCCL all-reduces are captured as graph nodes)."""

    def __init__(self, model, optimizer, x, y, M=None, warmup=3, eps=None):
        from optim import FusedAdamW
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("GraphedTrainStep needs optim.FusedAdamW (its step count can live on the device)")
        N.require_cuda(x, y)
        self.model, self.opt, self.M = model, optimizer, M
        self.x, self.y = x.detach().clone(), y.detach().clone()
        self.eps = eps.detach().clone() if eps is not None else None      # injected N(0,1) draws (parity tests only)
        self.counters = torch.zeros(2, device=x.device, dtype=torch.int32)
        host_step = max(g.get("step", 0) for g in optimizer.param_groups)
        self.counters[0] = host_step
        self.counters[1] = int(torch.initial_seed() & 0x7FFFFFFF)
        self._saved_sync = model.sync_scalars
        model.sync_scalars = False
        optimizer.use_device_step(self.counters[0:1])
        N.check(N.lib().pub_debug_pointer(b"seed_salt", C.c_void_p(self.counters[1:2].data_ptr())), "seed_salt")
        # Autograd binds every parameter's AccumulateGrad node to the stream that was current when the node was created,
        # and keeps the node while any graph references it.  model.{prior,posterior}_latent_space hold the encoders'
        # graph of the LAST step, so nodes created by earlier eager steps on the (legacy) default stream would survive
        # into the capture, where touching the legacy stream is an error: drop them, and warm up / capture on ONE stream.
        self._drop_graph_references()
        self.stream = torch.cuda.Stream()
        # eager warm-up on that stream: allocates workspaces / optimizer state, sets kernel attributes, builds NCCL
        # communicators -- nothing of that may happen for the first time during capture
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for i in range(max(1, warmup)):
                self._body()
                _dbg(f"warm-up step {i} enqueued")
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        _dbg("warm-up done")
        self.graph = None
        self.recapture()
        _dbg(f"captured {self.launches_per_step} launches")

    def _drop_graph_references(self):
        import gc
        self.model.prior_latent_space = None
        self.model.posterior_latent_space = None
        self.out = None
        gc.collect()

    def _body(self):
        N.check(N.lib().pub_advance_counters(N.ptr(self.counters), N.stream()), "pub_advance_counters")
        self.opt.zero_grad(set_to_none=True)
        kw = {"eps": self.eps} if self.eps is not None else {}
        out = self.model.elbo(self.x, self.y, None, M=self.M, **kw) if self.M is not None else self.model.elbo(self.x, self.y, None, **kw)
        out[0].backward()
        self.opt.step()
        return out

    def recapture(self):
        self._drop_graph_references()
        self.opt.zero_grad(set_to_none=True)
        l0 = N.lib().pub_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = self._body()
        except Exception as ex:
            self.graph = None
            raise N.NativeError(
                "CUDA-graph capture of the training step failed.  The usual cause: a tensor from an earlier EAGER step (the "
                "last loss, a kept elbo() output) still references that step's autograd graph, whose AccumulateGrad nodes "
                "are bound to the default stream -- which must not be touched during capture.  Drop those references "
                f"(`del loss`) before constructing GraphedTrainStep.  Original error: {ex!r}") from ex
        self.launches_per_step = int(N.lib().pub_launch_count() - l0)

    def __call__(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        _dbg("replay enqueued")
        return self.out

    def close(self):
        """Back to eager stepping: host step count, no device salt."""
        steps = int(self.counters[0].item())
        for g in self.opt.param_groups:
            g["step"] = steps
        self.opt.use_device_step(None)
        N.lib().pub_debug_pointer(b"seed_salt", C.c_void_p(0))
        self.model.sync_scalars = self._saved_sync
        self.graph = None

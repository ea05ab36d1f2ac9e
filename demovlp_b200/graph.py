"""Whole-step CUDA graph for the loss path (optional helper; the eager modules stay the drop-in).

``GraphedLossStep`` captures ONE graph holding the forward of a loss callable, ``loss.backward()`` and the
stacking of the returned scalars, on static input buffers: a replay costs one launch instead of ~30, which is
what the step is bound by at per-GPU batch 128 (the kernels total ~0.5 ms).  The caller writes the next batch
into ``step.inputs[...]`` (e.g. ``copy_(pinned, non_blocking=True)``), calls ``replay()`` and reads
``step.grads[...]`` / ``step.losses`` (device tensors, valid until the next replay).
Shapes, dtypes and the set of tensors that require grad are frozen at capture time.
"""
from __future__ import annotations

import torch

from . import _lib


class GraphedLossStep:
    def __init__(self, fn, inputs, grad_keys, warmup=3):
        """fn(dict of tensors) -> tuple of scalar tensors, the first being the loss to back-propagate."""
        self.fn, self.grad_keys = fn, tuple(grad_keys)
        self.inputs = {k: v.detach().clone() for k, v in inputs.items()}
        for k in self.grad_keys:
            self.inputs[k].requires_grad_(True)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside the capture: lazy initialisation, autotuned attributes
            for _ in range(warmup):
                self._clear_grads()
                self.fn(self.inputs)[0].backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._clear_grads()
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = self.fn(self.inputs)
            out[0].backward()
            self.losses = torch.stack([o.detach().reshape(()).float() for o in out])
        self.launches_per_replay = _lib.launch_count() - n0      # library kernels captured (torch glue not counted)
        self.grads = {k: self.inputs[k].grad for k in self.grad_keys}

    def _clear_grads(self):
        for k in self.grad_keys:
            self.inputs[k].grad = None

    def replay(self):
        self.graph.replay()
        return self.losses

"""CUDA-graph replay of the cloze step for a fixed batch shape.

The step launches ~900 kernels through the C ABI; at ~95 ms per step the launch gaps are
worth 4-5 %.  `GraphedStep` captures mask -> embed -> block stack -> heads -> backward once
(torch.cuda.graph: every libgraft launch goes to the capturing stream like any other
kernel) and replays it per batch.  The optimiser tail (gradient all-reduce, clip, AdamW)
stays outside the graph: AdamW's bias correction changes every step and NCCL keeps its own
stream semantics.  The mask draw (torch.rand) is graph-safe: PyTorch re-seeds the Philox
offset per replay.
"""
from __future__ import annotations

import torch

from . import lib as L


class GraphedStep:
    def __init__(self, model, task_weights, example_batch, warmup=2):
        """`example_batch`: dict of device tensors with the shapes/dtypes every later batch has
        (the trainer's flat [B*S] arrays).  Parameters must already have their .grad buffers
        (optim.FlatAdamW points them at its flat gradient buffer)."""
        self.model, self.tw = model, [float(w) for w in task_weights]
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.shapes = {k: v.shape for k, v in self.static.items()}
        self.graph = None
        self.losses = None
        self.launches = 0
        # the graph bakes in the raw pointers of the workspaces its kernels borrowed (split-K partials, sort space,
        # reduction partials): they live in a scope this object owns, so no later eager call can replace or free them
        self.scratch = L.ScratchScope()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with self.scratch, torch.cuda.stream(side):  # warm-up on a side stream: allocator pools, lazy caches, function attributes
            for _ in range(warmup):
                self._load(example_batch)
                self._eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._load(example_batch)
        g = torch.cuda.CUDAGraph()
        before = L.LAUNCHES
        with self.scratch, torch.cuda.graph(g):
            self.losses = self._eager()
        self.scratch.frozen = True
        self.launches = L.LAUNCHES - before
        self.graph = g

    def _load(self, batch):
        for k, v in batch.items():
            # the step reshapes and masks its batch in place: restore the flat views' contents
            self.static[k].view(-1).copy_(v.reshape(-1), non_blocking=True)

    def _eager(self):
        d = {k: v.view(self.shapes[k]) for k, v in self.static.items()}
        losses = self.model(d, False)
        total = sum(l * w for l, w in zip(losses, self.tw))
        total.backward()
        return torch.stack([l.detach() for l in losses])

    def __call__(self, batch):
        """batch: dict of host (pinned) or device tensors.  Returns the 4 losses (device tensor,
        overwritten by the next call); parameter .grad buffers hold this batch's gradient sum
        on top of what they held (zero them between optimiser steps, as usual)."""
        self._load(batch)
        self.graph.replay()
        L.LAUNCHES += self.launches
        return self.losses

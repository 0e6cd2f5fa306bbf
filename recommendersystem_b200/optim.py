"""Step tail on flat buffers: gradient all-reduce, global-norm clip, fused AdamW.

Replaces the reference's per-step tail (transformer.py:272-276, 285-298: DDP
bucketed all-reduce, clip_grad_norm_(1.0), optim.AdamW(fused=True, betas=(0.9,0.95),
weight_decay 0.1 on dim>=2 tensors)) with three libgraft launches over one
contiguous fp32 parameter buffer and one gradient buffer:
  ncclAllReduce(flat grad) -> sum of squares -> AdamW (decay / no-decay halves).
Parameters are re-pointed at views of the flat buffer, `.grad` at views of the
flat gradient buffer, so autograd accumulates in place and nothing is copied.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import lib as L


# loss scales of the four tasks, order [0.watch, 0.rating, 1.watch, 1.rating] (transformer.py:381-390)
TASK_LOSS_SCALE = (4.618602403897067, 1.1958987168236102, 2.5443243303769867, 1.0527565486045412)


def make_task_weights(finetune_medium=None, finetune_metric=None):
    """`make_task_weights` of the reference trainer (transformer.py:379-406): per-task weights of the step's loss sum,
    order [0.watch, 0.rating, 1.watch, 1.rating].  Pre-train: anime 4x manga, watch 4x rating; finetune: one
    (medium, metric) pair carries all the weight.  Each weight is normalised to sum 1, then divided by the task's scale."""
    metrics = ("watch", "rating")
    balance = []
    for medium in (0, 1):
        wm = (0.25, 1.0)[medium] if finetune_medium is None else float(medium == finetune_medium)
        for metric in metrics:
            wk = (1.0, 0.25)[metrics.index(metric)] if finetune_metric is None else float(metric == finetune_metric)
            balance.append(wm * wk)
    total = sum(balance)
    return [(b / total) / s for b, s in zip(balance, TASK_LOSS_SCALE)]


def wsd_lr_factor(step, warmup_steps, total_steps, decay_ratio=0.1, final_ratio=0.1):
    """Warm-up / stable / decay multiplier of the reference's pre-train schedule (transformer.py:310-328
    WSDScheduler, used as LambdaLR at :344-347: warm-up 2000 steps, last 10 % decays linearly to 0.1x).
    `step` counts scheduler.step() calls, as LambdaLR passes it."""
    decay_steps = int(total_steps * decay_ratio)
    stable_steps = total_steps - warmup_steps - decay_steps
    assert stable_steps >= 0
    s = max(0, min(int(step), total_steps))
    if s <= warmup_steps:
        return s / max(1, warmup_steps)
    if s <= warmup_steps + stable_steps:
        return 1.0
    return 1.0 - (1.0 - final_ratio) * (s - warmup_steps - stable_steps) / max(1, decay_steps)


class FlatAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1, max_grad_norm=1.0, graft_comm=None):
        """graft_comm: a comm.GraftComm -> the gradient all-reduce goes through the C ABI (graft_allreduce_f32, the library's
        own NCCL communicator) instead of torch.distributed."""
        self.graft_comm = graft_comm
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        decay = [(n, p) for n, p in named if p.dim() >= 2]      # transformer.py:288-289
        nodecay = [(n, p) for n, p in named if p.dim() < 2]
        self.params = [p for _, p in decay + nodecay]
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]  # keep every view 16-byte aligned
        total = sum(sizes)
        self.n_decay = sum(sizes[: len(decay)])
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p, sz in zip(self.params, sizes):
            view = self.flat_p[off: off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off: off + p.numel()].view_as(p)
            off += sz
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.steps = 0
        self._sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scale = torch.ones(1, dtype=torch.float32, device=dev)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def zero_grad(self):
        self.flat_g.zero_()

    def all_reduce_grads(self):
        """Data-parallel reduction of the dense gradient: ONE collective on the flat buffer
        (the reference's DDP reducer, transformer.py:272/678-682, moves the same bytes in 25 MB
        buckets).  The sum is reduced; the 1/world of DDP's mean is folded into the clip scale
        that AdamW applies anyway, so no extra pass over the buffer is made."""
        if self.world > 1:
            if self.graft_comm is not None:
                self.graft_comm.all_reduce_(self.flat_g)
            else:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)

    def clip_scale(self, sumsq):
        """Scale applied to the (summed) gradient: DDP mean x clip_grad_norm_(max_norm)
        (transformer.py:273): min(1, max_norm / (||g_mean|| + 1e-6)) / world, on device."""
        norm_mean = sumsq.sqrt() / self.world
        return torch.clamp(self.max_norm / (norm_mean + 1e-6), max=1.0) / self.world

    def step(self, lr=None):
        self.steps += 1
        # graft_adamw updates the parameters through raw pointers, which does not bump tensor._version: tell the
        # model-side caches (fused item table, bf16 inference weights) that every weight changed
        from .model import ItemEmbedding
        ItemEmbedding._weights_version += 1
        lr = self.lr if lr is None else lr
        self.all_reduce_grads()
        # clip_grad_norm_(params, max_norm): scale = min(1, max_norm / (norm + 1e-6)), on device
        self._sq.zero_()
        L.sumsq(self.flat_g, self._sq)
        self._scale.copy_(self.clip_scale(self._sq))
        b1, b2 = self.betas
        nd = self.n_decay
        L.adamw(self.flat_p[:nd], self.flat_g[:nd], self.m[:nd], self.v[:nd], lr, b1, b2, self.eps, self.wd, self.steps, self._scale)
        if nd < self.flat_p.numel():
            L.adamw(self.flat_p[nd:], self.flat_g[nd:], self.m[nd:], self.v[nd:], lr, b1, b2, self.eps, 0.0, self.steps, self._scale)

    def set_schedule(self, warmup_steps, total_steps, decay_ratio=0.1, final_ratio=0.1):
        """Pre-train WSD schedule (transformer.py:344-347); finetune keeps the constant factor 1 (:331-332)."""
        self._sched = (warmup_steps, total_steps, decay_ratio, final_ratio)
        self._sched_step = 0

    def scheduled_step(self):
        """optimizer.step(); scheduler.step() of transformer.py:274-276: the n-th call uses the factor of step n-1."""
        f = wsd_lr_factor(self._sched_step, *self._sched) if getattr(self, "_sched", None) else 1.0
        self.step(lr=self.lr * f)
        if getattr(self, "_sched", None):
            self._sched_step += 1
        return f

    def grad_norm(self):
        """Global L2 norm of the rank-averaged gradient of the last step (device tensor)."""
        return self._sq.sqrt() / self.world

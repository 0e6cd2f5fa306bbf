"""Collective plumbing for the multi-GPU paths (torch.distributed; NCCL over NVLink on the box).

The step has exactly one exchange in data-parallel mode (the flat gradient all-reduce,
optim.FlatAdamW) and three small ones in vocab-parallel mode (vocab_parallel notes in
model.py): gather of masked hidden rows, exchange of per-row softmax statistics,
reduce-scatter of the hidden-row gradient.  With the NCCL backend these map 1:1 onto
ncclAllGather / ncclAllToAll / ncclReduceScatter; with gloo (CPU tests, or two test
processes sharing one GPU) the same results are produced from all_gather / all_reduce on
host copies.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def _native(t, group=None):
    return dist.get_backend(group) == "nccl" and t.is_cuda


def all_gather_rows(t, group=None):
    """[n, ...] on every rank -> [P*n, ...], rank-major."""
    P = world(group)
    if P == 1:
        return t
    t = t.contiguous()
    out = torch.empty((P * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if _native(t, group):
        dist.all_gather_into_tensor(out, t, group=group)
        return out
    host = t.cpu()
    if host.dtype == torch.bfloat16:  # gloo has no bf16: move the raw bytes
        host = host.view(torch.uint8)
    parts = [torch.empty_like(host) for _ in range(P)]
    dist.all_gather(parts, host, group=group)
    cat = torch.cat(parts, 0)
    if t.dtype == torch.bfloat16:
        cat = cat.view(torch.bfloat16)
    out.copy_(cat)
    return out


def reduce_scatter_rows(t, group=None):
    """[P*n, ...] partial sums on every rank -> this rank's [n, ...] slice of the total (fp32)."""
    P = world(group)
    if P == 1:
        return t
    n = t.shape[0] // P
    t = t.contiguous()
    if _native(t, group):
        out = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.SUM, group=group)
        return out
    host = t.cpu()
    dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
    r = rank(group)
    return host[r * n:(r + 1) * n].to(t.device)


def all_to_all_rows(t, group=None):
    """t[q] is what this rank holds for rank q ([P, ...]); returns out[p] = what rank p held for this rank."""
    P = world(group)
    if P == 1:
        return t
    t = t.contiguous()
    if _native(t, group):
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=group)
        return out
    host = t.cpu()
    parts = [torch.empty_like(host) for _ in range(P)]
    dist.all_gather(parts, host, group=group)
    r = rank(group)
    return torch.stack([p[r] for p in parts]).to(t.device)

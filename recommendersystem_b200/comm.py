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


class GraftComm:
    """Library-owned NCCL communicator (include/graft.h: graft_comm_*): the data-parallel all-reduce of the flat gradient
    buffer issued through the C ABI on any stream.  Rank 0 draws the NCCL unique id; it reaches the other ranks through the
    already initialised torch.distributed group (any backend)."""

    def __init__(self, device=None, group=None):
        import ctypes as C
        from . import lib as L
        self._lib = L.load()
        self.rank, self.world = rank(group), world(group)
        if device is not None:
            torch.cuda.set_device(device)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            L._check(self._lib.graft_comm_unique_id(buf), "graft_comm_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            t = ident.cuda() if on_gpu else ident
            dist.broadcast(t, src=0, group=group)
            ident = t.cpu()
        raw = (C.c_ubyte * 128)(*ident.tolist())
        self._comm = C.c_void_p()
        L._check(self._lib.graft_comm_create(raw, self.rank, self.world, C.byref(self._comm)), "graft_comm_create")

    def all_reduce_(self, t, stream=None):
        """In-place SUM over the ranks of a contiguous fp32 CUDA tensor, stream-ordered on `stream` (default: current)."""
        import ctypes as C
        from . import lib as L
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        st = L.stream() if stream is None else L.vp(stream.cuda_stream)
        L._check(self._lib.graft_allreduce_f32(self._comm, L.vp(t.data_ptr()), C.c_int64(t.numel()), st), "graft_allreduce_f32")
        return t

    def close(self):
        if self._comm:
            self._lib.graft_comm_destroy(self._comm)
            self._comm = None

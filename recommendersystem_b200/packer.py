"""Host mirror of the packing / batching either side of the step.

Mirrors, with the same names and argument meaning:
  * PretrainDataset.get_index_permutation / block_shuffle / window slicing
    (/root/reference/notebooks/Training/transformer.py:53-98)
  * the shard-by-rank rule (transformer.py:46-51)
  * concat (/root/reference/notebooks/Training/transformer.jl:144-163)
The index arithmetic runs in libgraft's plain-C host entry points
(graft_jagged_offsets_host, graft_block_permutation_host, graft_concat_pad_host:
the functions the Julia packer binds with ccall, integration/graft.jl), not in
Python loops.  No CUDA context is needed for anything in this file.
"""
from __future__ import annotations

import numpy as np

from . import lib as L


def jagged_offsets(userid):
    """Run boundaries of equal userid: offsets[0]=0 ... offsets[-1]=len (transformer.py:54)."""
    return L.jagged_offsets_host(userid)


def index_permutation(userid, permutation=np.random.permutation):
    """transformer.py:53-68: permutes whole user blocks; `permutation(n)` supplies the block order."""
    off = jagged_offsets(userid)
    order = np.asarray(permutation(len(off) - 1), dtype=np.int64)
    return L.block_permutation_host(off, order)


def block_shuffle(d, permutation=np.random.permutation):
    """transformer.py:70-73 (in place on the dict of flat arrays)."""
    p = index_permutation(d["userid"], permutation)
    for k in d:
        d[k] = d[k][p]
    return d


def windows(d, tokens_per_batch):
    """transformer.py:90-98: consecutive windows of tokens_per_batch positions."""
    n = len(d["userid"])
    assert n % tokens_per_batch == 0, "file length must be a multiple of the batch (transformer.py:91)"
    for i in range(0, n, tokens_per_batch):
        yield {k: v[i: i + tokens_per_batch] for k, v in d.items()}


def shards_for_rank(shards, local_rank, local_world_size):
    """transformer.py:46-51: shard directory i belongs to rank i % world."""
    shards = sorted(shards)
    assert len(shards) % local_world_size == 0
    return [s for i, s in enumerate(shards) if i % local_world_size == local_rank]


def concat(per_user, batch_size):
    """transformer.jl:144-163: every key back to back over users, zero padded to a multiple of batch_size."""
    n = sum(len(u["userid"]) for u in per_user)
    total = n + (-n) % batch_size
    return {k: L.concat_pad_host([u[k] for u in per_user], total) for k in per_user[0]}

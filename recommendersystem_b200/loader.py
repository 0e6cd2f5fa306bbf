"""Input pipeline of the pre-train loop: shard files -> user-block shuffle -> windows ->
pinned staging -> asynchronous H2D.

Mirrors PretrainDataset + collate + to_device
(/root/reference/notebooks/Training/transformer.py:37-98, 167-184, 633-646):
  * shard directory i belongs to local rank i % local_world_size           (:46-51)
  * per file: read the 27 flat arrays, permute whole user runs             (:53-72)
  * yield consecutive windows of tokens_per_batch positions                (:90-98)
  * one batch = dict of flat tensors on the GPU                            (:167-184)
The reference does this with 8 DataLoader worker processes per GPU, a Python loop per
user block and per-key fancy indexing.  Here the index arithmetic is the C entry points
the Julia packer also binds (graft_jagged_offsets_host, graft_block_permutation_host), a
background thread prepares the next file while the current one is consumed, windows are
written into two rotating sets of pinned host buffers, and the H2D copies run on their own
stream so that they overlap the previous step (at 0.7 M interactions/s one GPU consumes
~80 MB/s of packed arrays: 112 B per interaction).

Shard files: the reference writes HDF5 + blosc (transformer.jl:228-232); h5py/hdf5plugin are
not in this image, so `.npz` files with the same 27 keys are read as well (tests, bench).
"""
from __future__ import annotations

import glob
import os
import queue
import threading

import numpy as np
import torch

from . import packer


def read_shard_file(fn):
    """One packed file -> dict of flat numpy arrays (transformer.py:84-88)."""
    if fn.endswith(".npz"):
        with np.load(fn) as z:
            return {k: z[k] for k in z.files}
    import h5py  # the reference's format; optional dependency
    try:
        import hdf5plugin  # noqa: F401  (registers the blosc filter)
    except ImportError:
        pass
    with h5py.File(fn, "r") as f:
        return {k: f[k][:] for k in f}


class PretrainLoader:
    """Iterating yields dicts of flat device tensors of `tokens_per_batch` positions."""

    def __init__(self, datadir, local_rank, local_world_size, tokens_per_batch, device=None, seed=None, shuffle_files=True,
                 prefetch_files=1):
        self.batch_size = int(tokens_per_batch)
        shards = sorted(glob.glob(os.path.join(datadir, "*", "")))
        self.fns = []
        for s in packer.shards_for_rank(shards, local_rank, local_world_size):
            self.fns.extend(sorted(glob.glob(os.path.join(s, "*.h5")) + glob.glob(os.path.join(s, "*.npz"))))
        self.device = torch.device(device) if device is not None else None
        self.rng = np.random.default_rng(seed)
        self.shuffle_files = shuffle_files
        self.prefetch_files = prefetch_files
        self._pinned = None
        self._slot = 0
        self._stream = None
        self._busy = [None, None]

    # ---- host side -------------------------------------------------------------------
    def _prepare(self, fn):
        d = read_shard_file(fn)
        packer.block_shuffle(d, self.rng.permutation)          # whole users move, transformer.py:70-73
        assert len(d["userid"]) % self.batch_size == 0, fn     # transformer.py:91
        return d

    def host_batches(self):
        """Windows as numpy views, files prepared one ahead by a background thread."""
        fns = list(self.fns)
        if self.shuffle_files:
            self.rng.shuffle(fns)                               # transformer.py:83
        q = queue.Queue(maxsize=max(1, self.prefetch_files))

        def work():
            try:
                for fn in fns:
                    q.put(self._prepare(fn))
                q.put(None)
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        while True:
            d = q.get()
            if d is None:
                break
            if isinstance(d, BaseException):
                raise d
            yield from packer.windows(d, self.batch_size)
        t.join()

    # ---- device side -----------------------------------------------------------------
    def _stage(self, window):
        """numpy window -> pinned buffers of the current slot -> async H2D on the copy stream."""
        if self._pinned is None:
            self._pinned = [{k: torch.empty(v.shape, dtype=torch.from_numpy(v[:0]).dtype).pin_memory() for k, v in window.items()}
                            for _ in range(2)]
            self._stream = torch.cuda.Stream(device=self.device)
        slot = self._slot
        self._slot ^= 1
        if self._busy[slot] is not None:
            self._busy[slot].synchronize()                      # the copies that last used these pinned buffers are done
        host = self._pinned[slot]
        for k, v in window.items():
            host[k].numpy()[...] = v
        with torch.cuda.stream(self._stream):
            dev = {k: t.to(self.device, non_blocking=True) for k, t in host.items()}
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._busy[slot] = ev
        return dev, ev

    def __iter__(self):
        if self.device is None or self.device.type != "cuda":
            for w in self.host_batches():                       # CPU consumers (tests): plain tensors
                yield {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}
            return
        pending = None
        for w in self.host_batches():
            nxt = self._stage(w)                                # copy of batch i+1 overlaps the step on batch i
            if pending is not None:
                dev, ev = pending
                torch.cuda.current_stream(self.device).wait_event(ev)
                for t in dev.values():
                    t.record_stream(torch.cuda.current_stream(self.device))
                yield dev
            pending = nxt
        if pending is not None:
            dev, ev = pending
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in dev.values():
                t.record_stream(torch.cuda.current_stream(self.device))
            yield dev


def write_shards(datadir, per_user, n_shards, users_per_file, batch_size, seed=0):
    """Test/bench helper restating save_data (transformer.jl:202-240) for synthetic users: users
    round-robin over `n_shards` directories, files of `users_per_file` users, each file concat +
    zero padded to a multiple of batch_size, userid = 1-based index inside its file (:225)."""
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(per_user))
    for s in range(n_shards):
        dest = os.path.join(datadir, str(s + 1))
        os.makedirs(dest, exist_ok=True)
        mine = [per_user[i] for j, i in enumerate(order) if j % n_shards == s]
        for p, start in enumerate(range(0, len(mine), users_per_file)):
            chunk = mine[start:start + users_per_file]
            for i, d in enumerate(chunk):
                d = dict(d)
                d["userid"] = np.full(len(d["userid"]), i + 1, np.int32)
                chunk[i] = d
            np.savez(os.path.join(dest, f"{p + 1}.npz"), **packer.concat(chunk, batch_size))


class FinetuneDataset:
    """Mirror of the reference's FinetuneDataset + collate + to_device (transformer.py:101-184) for the finetune step: files
    hold [users, max_seq_len] matrices (Finetune/transformer.jl:135-166); only users with a target of the finetuned medium
    are kept (:136-143); training takes every `checkpoints_per_epoch`-th such user of the current partition, shuffles them and
    pads the list to a multiple of batch_size by re-drawing (:144-152); batches are dicts of [batch_size, max_seq_len]
    tensors, `.position` as int64 on the device (:179-183).  The row filter is one vectorised reduction instead of the
    reference's Python loop over users; iteration order and draws follow the reference (np.random, as seeded by the caller)."""

    def __init__(self, datadir, local_rank, local_world_size, batch_size, shuffle, finetune_medium, device=None, checkpoints_per_epoch=4):
        self.batch_size, self.shuffle, self.medium = int(batch_size), bool(shuffle), int(finetune_medium)
        shards = sorted(glob.glob(os.path.join(datadir, "*", "")))
        self.fns = []
        for s in packer.shards_for_rank(shards, local_rank, local_world_size):
            self.fns.extend(glob.glob(os.path.join(s, "*.h5")) + glob.glob(os.path.join(s, "*.npz")))
        self.partition = [0, checkpoints_per_epoch]
        self.device = torch.device(device) if device is not None else None

    def rows_with_targets(self, d):
        """transformer.py:136-143: users whose finetuned medium has a watch or rating weight."""
        w = d[f"{self.medium}.watch.weight"].sum(axis=1) > 0
        r = d[f"{self.medium}.rating.weight"].sum(axis=1) > 0
        return np.flatnonzero(w | r).tolist()

    def __iter__(self):
        for fn in self.fns:
            d = read_shard_file(fn)
            idxs = self.rows_with_targets(d)
            if self.shuffle:
                idxs = [x for i, x in enumerate(idxs) if i % self.partition[1] == self.partition[0]]
                np.random.shuffle(idxs)
                while len(idxs) % self.batch_size != 0:
                    idxs.append(np.random.choice(idxs))
            for i in range(0, len(idxs), self.batch_size):
                idx = idxs[i: i + self.batch_size]
                batch = {k: torch.as_tensor(np.ascontiguousarray(v[idx, :])) for k, v in d.items()}
                if self.device is not None:
                    batch = {k: (v.pin_memory() if self.device.type == "cuda" else v).to(self.device, non_blocking=True) for k, v in batch.items()}
                    batch = {k: (v.to(torch.int64) if k.endswith(".position") else v) for k, v in batch.items()}
                yield batch
        self.partition[0] = (self.partition[0] + 1) % self.partition[1]

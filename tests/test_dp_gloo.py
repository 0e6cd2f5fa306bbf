"""World-size-2 checks of the data-parallel host logic on CPU (gloo): the shard-by-rank
rule (transformer.py:46-51), the single flat-buffer gradient reduction that replaces the
DDP reducer (transformer.py:272, 678-682) and the folded mean/clip scale (transformer.py:273).
The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from recommendersystem_b200 import packer
from recommendersystem_b200.optim import FlatAdamW


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical replicas, as DDP broadcasts them
        model = nn.Sequential(nn.Linear(6, 5), nn.Linear(5, 3))
        opt = FlatAdamW(model, lr=1e-3)
        assert opt.world == world
        # parameters and grads are views of the flat buffers
        for p in model.parameters():
            assert p.data.untyped_storage().data_ptr() == opt.flat_p.untyped_storage().data_ptr()
            assert p.grad.untyped_storage().data_ptr() == opt.flat_g.untyped_storage().data_ptr()
        # decay group first (dim >= 2), as transformer.py:288-289 splits them
        assert opt.n_decay == 32 + 16  # 30 and 15 weights, each view padded to 16 bytes
        x = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 + rank))
        model(x).square().sum().backward()
        local = opt.flat_g.clone()
        opt.all_reduce_grads()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        total = torch.stack(gathered).sum(0)
        assert torch.allclose(opt.flat_g, total, rtol=1e-6, atol=1e-7)
        # folded scale == DDP mean followed by clip_grad_norm_(1.0)
        mean = total / world
        ref_scale = min(1.0, 1.0 / (float(mean.norm()) + 1e-6))
        sc = opt.clip_scale((opt.flat_g ** 2).sum())
        assert abs(float(sc) * world - ref_scale) < 1e-6 * ref_scale
        assert torch.allclose(opt.flat_g * sc, mean * ref_scale, rtol=1e-5, atol=1e-8)
        # shard rule: disjoint cover of the 8 shard directories the packer writes (transformer.jl:21)
        shards = [f"/data/transformer/training/{i}/" for i in range(1, 9)]
        mine = packer.shards_for_rank(shards, rank, world)
        names = [None] * world
        dist.all_gather_object(names, mine)
        flat = sorted(sum(names, []))
        assert flat == sorted(shards) and len(set(flat)) == 8
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def test_block_shuffle_moves_whole_users():
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 30, 40)
    u = np.repeat(np.arange(1, 41, dtype=np.int32), lens)
    d = {"userid": u.copy(), "matchedid": np.arange(len(u), dtype=np.int32)}
    order = rng.permutation(40)
    packer.block_shuffle(d, lambda n: order)
    # the literal reference algorithm, transformer.py:53-68
    change = np.where(u[:-1] != u[1:])[0] + 1
    blocks = np.split(np.arange(len(u)), change)
    ref = np.concatenate([blocks[b] for b in order])
    assert np.array_equal(d["matchedid"], ref) and np.array_equal(d["userid"], u[ref])
    w = list(packer.windows({"userid": np.zeros(64, np.int32)}, 16))
    assert len(w) == 4 and all(len(x["userid"]) == 16 for x in w)


def test_wsd_schedule_matches_reference_formula():
    """transformer.py:310-328 restated literally next to optim.wsd_lr_factor."""
    from recommendersystem_b200.optim import wsd_lr_factor

    class WSD:  # the reference class, verbatim semantics
        def __init__(self, warmup_steps, total_steps, decay_ratio, final_ratio):
            self.warmup_steps, self.total_steps, self.final_ratio = warmup_steps, total_steps, final_ratio
            self.decay_steps = int(total_steps * decay_ratio)
            self.stable_steps = total_steps - warmup_steps - self.decay_steps

        def __call__(self, step):
            s = max(0, min(int(step), self.total_steps))
            if s <= self.warmup_steps:
                return s / max(1, self.warmup_steps)
            if s <= (self.warmup_steps + self.stable_steps):
                return 1.0
            return 1.0 - (1.0 - self.final_ratio) * (s - self.warmup_steps - self.stable_steps) / max(1, self.decay_steps)

    for w, t in ((2000, 50000), (10, 37), (0, 5)):
        ref = WSD(w, t, 0.1, 0.1)
        for s in list(range(0, min(t, 60))) + [t // 2, t - 3, t - 1, t, t + 5]:
            assert wsd_lr_factor(s, w, t) == ref(s), (w, t, s)


def test_make_task_weights_matches_reference_formula():
    """optim.make_task_weights (the product's mirror of transformer.py:379-406) == the oracle's restatement, for the
    pre-train weights and every finetune (medium, metric) choice."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.optim import make_task_weights
    for fm in (None, 0, 1):
        for fk in (None, "watch", "rating"):
            a, b = make_task_weights(fm, fk), default_task_weights(fm, fk)
            assert len(a) == 4 and all(abs(x - y) <= 1e-15 * max(1.0, abs(y)) for x, y in zip(a, b)), (fm, fk, a, b)
    w = make_task_weights()
    assert abs(sum(x * s for x, s in zip(w, (4.618602403897067, 1.1958987168236102, 2.5443243303769867, 1.0527565486045412))) - 1.0) < 1e-12

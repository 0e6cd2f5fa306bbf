"""graft_allreduce_f32 (include/graft.h): the library-owned NCCL communicator sums a flat fp32 buffer over
the ranks exactly like torch.distributed's all_reduce, on the current stream and on a side stream.  Needs two GPUs (one rank
per GPU); on a one-GPU box only the single-rank path runs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        from recommendersystem_b200.comm import GraftComm
        gc = GraftComm(device=rank)
        g = torch.Generator().manual_seed(rank)
        x = torch.randn(3_000_001, generator=g).cuda()
        ref = x.clone()
        dist.all_reduce(ref)
        got = gc.all_reduce_(x.clone())
        assert torch.equal(got, ref)                      # same ring, same order: bit-identical
        side = torch.cuda.Stream()
        y = x.clone()
        side.wait_stream(torch.cuda.current_stream())
        gc.all_reduce_(y, stream=side)
        torch.cuda.current_stream().wait_stream(side)
        assert torch.allclose(y, ref, rtol=1e-6, atol=1e-6)
        gc.close()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_single_rank_communicator_is_identity():
    from recommendersystem_b200.comm import GraftComm
    gc = GraftComm(device=0)
    x = torch.arange(1000, dtype=torch.float32, device="cuda:0")
    assert torch.equal(gc.all_reduce_(x.clone()), x)
    gc.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (one rank per GPU over NCCL / NVLink)")
def test_graft_allreduce_matches_torch_distributed_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert list(out) == [1] * world

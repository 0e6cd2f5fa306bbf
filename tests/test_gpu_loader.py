"""Loader on the GPU: pinned double-buffered staging + copy-stream H2D delivers exactly the
windows the host path yields, and a train step runs on them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from recommendersystem_b200 import loader, packer, synthetic  # noqa: E402


def test_device_batches_equal_host_batches_and_feed_a_step(tmp_path):
    cfg, rows = synthetic.make_config("tiny")
    S = cfg["max_sequence_length"]
    smp = synthetic.HistorySampler(cfg, seed=3, len_median=20.0, len_clip=(3, 60))
    flat, _ = smp.pack(60, rows * S)
    off = packer.jagged_offsets(flat["userid"])
    users = [{k: v[a:b] for k, v in flat.items()} for a, b in zip(off[:-1], off[1:]) if flat["userid"][a] > 0]
    loader.write_shards(str(tmp_path), users, n_shards=2, users_per_file=16, batch_size=rows * S, seed=2)
    host = list(loader.PretrainLoader(str(tmp_path), 0, 1, rows * S, seed=9))
    dev = list(loader.PretrainLoader(str(tmp_path), 0, 1, rows * S, device="cuda:0", seed=9))
    assert len(host) == len(dev) >= 2
    for a, b in zip(host, dev):
        for k in a:
            assert b[k].is_cuda and torch.equal(a[k], b[k].cpu()), k
    from recommendersystem_b200.model import RecommenderModel
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=synthetic.make_metadata(cfg))
    model = model.to("cuda:0")
    losses = model(dev[0], False)
    sum(losses).backward()
    assert all(np.isfinite(float(l)) for l in losses)

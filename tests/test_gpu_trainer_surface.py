"""The wrappers the reference trainer puts around the model: .to(rank), torch.compile, DDP,
no_sync gradient accumulation, autocast(bf16), named_parameters-driven AdamW groups, clip_grad_norm_,
reading d[...weight].sum() after the call (transformer.py:256-276, 285-298, 670-682) - run here as a
two-step single-process train loop on the drop-in."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def test_reference_trainer_loop_runs_on_the_drop_in(golden):
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.testing import batch_to_device, model_from_params
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        cfg = golden["config"]
        model = model_from_params(cfg, golden["param"], "cpu")
        model = model.to(0)                                            # transformer.py:670
        model = torch.compile(model)                                   # :677 (the step opts out of tracing)
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0])  # :678-682
        named = {n: p for n, p in model.named_parameters() if p.requires_grad}
        opt = torch.optim.AdamW([{"params": [p for p in named.values() if p.dim() >= 2], "weight_decay": 0.1},
                                 {"params": [p for p in named.values() if p.dim() < 2], "weight_decay": 0.0}],
                                lr=1e-3, betas=(0.9, 0.95), fused=True)  # :285-298
        tw = default_task_weights()
        accum = 2
        before = {n: p.detach().clone() for n, p in named.items()}
        seen = []
        for step in range(4):
            d = batch_to_device(golden["batch"], DEV)                  # flat [B*S] arrays, as to_device yields them
            with torch.amp.autocast("cuda", dtype=torch.bfloat16):     # :257
                tloss = model(d, False)                                # :259
            for m in (0, 1):                                           # :261: the trainer reads the post-mask weights
                for metric in ("watch", "rating"):
                    w = d[f"{m}.{metric}.weight"]                      # (DDP hands the module a shallow copy of the dict:
                    orig = float(golden["batch"][f"{m}.{metric}.weight"].sum())  # the tensors, masked in place, are shared)
                    assert w.numel() == golden["rows"] * cfg["max_sequence_length"] and 0 <= float(w.sum()) < orig
            loss = sum(tloss[i] * tw[i] for i in range(4)) / accum     # :264-267
            if (step + 1) % accum != 0:
                with model.no_sync():                                  # :268-271
                    loss.backward()
            else:
                loss.backward()                                        # :272
                gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)  # :273
                assert torch.isfinite(gn)
                opt.step(); opt.zero_grad(set_to_none=True)            # :274-275
            seen.append([float(x) for x in tloss])
        assert np.isfinite(np.array(seen)).all()
        moved = [n for n, p in named.items() if not torch.equal(p.detach(), before[n])]
        assert len(moved) == len(named), sorted(set(named) - set(moved))[:5]   # every trainable tensor was updated
        sd = model.module._orig_mod.state_dict()                       # the checkpoint path of :458
        assert set(k for k in sd if not k.startswith("watch_head.")) == set(golden["param"])
        with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):
            ev = model(batch_to_device(golden["batch"], DEV), True)    # evaluate_metrics, :220-227
        assert isinstance(ev[1], list) and len(ev[1]) == 3 and isinstance(ev[3], list)
    finally:
        dist.destroy_process_group()

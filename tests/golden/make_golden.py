"""Writes tests/golden/cloze_tiny.npz by running the REAL reference module.

Run inside the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The fixture pins oracle/cloze_oracle.py (tests/test_oracle_golden.py) and is the
golden vector the CUDA path is compared against on the GPU box.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
from oracle.cloze_oracle import default_task_weights  # noqa: E402
from recommendersystem_b200 import synthetic  # noqa: E402


def main(name="tiny", out="cloze_tiny.npz", param_seed=1234, data_seed=5, mask_seed=99):
    config, rows = synthetic.make_config(name)
    torch.set_num_threads(4)
    meta = synthetic.make_metadata(config, seed=3)
    model = rh.build_model(config, param_seed, metadata=meta)
    # the reference zero-inits biases, the mask rows and the time phases; perturb
    # them so the fixture also pins the gradient paths through those leaves
    g = torch.Generator().manual_seed(param_seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "periodic" in n or n.endswith("scale")):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        model.item_embedding.matchedid_embedding.embedding.weight[-1].normal_(0, 0.006, generator=g)
    batch = synthetic.make_batches(config, rows, 1, seed=data_seed, len_median=20.0, len_clip=(3, 80))[0]
    tw = default_task_weights()

    S = config["max_sequence_length"]
    torch.manual_seed(mask_seed)
    randval = torch.rand((rows, S))

    losses, dmask, grads = rh.run_step(model, batch, tw, mask_seed, autocast=False)
    losses_bf16, _, grads_bf16 = rh.run_step(model, batch, tw, mask_seed, autocast=True)
    with torch.no_grad():
        ev, _, _ = rh.run_step(model, batch, tw, mask_seed, evaluate=True)
    ev_flat = []
    for e in ev:
        ev_flat += [float(x) for x in e] if isinstance(e, list) else [float(e)]

    blob = {"config_json": np.frombuffer(json.dumps(config).encode(), dtype=np.uint8),
            "rows": np.int64(rows), "randval": randval.numpy(), "task_weights": np.array(tw),
            "losses": np.array([float(x) for x in losses]),
            "losses_bf16": np.array([float(x) for x in losses_bf16]),
            "eval_flat": np.array(ev_flat)}
    for k, v in batch.items():
        blob["batch/" + k] = v
    for k, v in dmask.items():
        blob["masked/" + k] = v.reshape(-1).numpy()
    for k, v in model.state_dict().items():
        if k.startswith("watch_head."):
            continue  # aliased duplicates of item_embedding.* (transformer.model.py:354)
        blob["param/" + k] = v.numpy()
    for k, v in grads.items():
        if k.startswith("watch_head."):
            continue
        blob["grad/" + k] = v.numpy()
        blob["grad_bf16/" + k] = grads_bf16[k].float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), out)
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; losses", blob["losses"], "bf16", blob["losses_bf16"])


if __name__ == "__main__":
    main()

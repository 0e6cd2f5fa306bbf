"""Writes tests/golden/finetune_tiny.npz by running the REAL reference module in finetune
mode (config finetune=True: LoRA r=8 on q,v, everything else frozen, item table fused at
load time; transformer.model.py:104-133, 235-271, 361-371, 418-435).  Build container only:
    python tests/golden/make_golden_finetune.py
Two passes: eval (dropout off) and train (nn.Dropout(0.1) on the LoRA inputs).  The train
pass's dropout masks are reproduced by drawing F.dropout(ones) from the same seed in the
reference's order (block by block, q then v) - asserted below against the oracle - and are
stored in the fixture so the CUDA path can be fed the same masks.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cloze_oracle as oc  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402
from recommendersystem_b200 import synthetic  # noqa: E402


def run(model, batch, tw, autocast, seed=None):
    d = rh.torch_batch(batch)
    model.zero_grad(set_to_none=True)
    if seed is not None:
        torch.manual_seed(seed)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        losses = model(d, False)
        loss = sum(losses[i] * tw[i] for i in range(4))
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return [float(l) for l in losses], grads, d


def main(out="finetune_tiny.npz", metric="watch"):
    cfg, rows = synthetic.make_config("tiny")
    torch.set_num_threads(4)
    meta = synthetic.make_metadata(cfg, seed=3)
    pre = rh.build_model(cfg, 1234, metadata=meta)          # the pre-trained checkpoint
    g = torch.Generator().manual_seed(1235)
    with torch.no_grad():
        for n, p in pre.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "periodic" in n or n.endswith("scale")):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        pre.item_embedding.matchedid_embedding.embedding.weight[-1].normal_(0, 0.006, generator=g)
    ckpt = {k: v.clone() for k, v in pre.state_dict().items()}

    cfg_ft = dict(cfg, finetune=True, finetune_metric=metric)
    ns = rh.load_namespace("cpu")
    torch.manual_seed(77)
    ft = ns["RecommenderModel"](cfg_ft)
    ft.load_state_dict(ckpt, strict=False)                  # transformer.py:650-656; the post-hook fuses the table
    assert ft.item_embedding.fused
    trainable = sorted(n for n, p in ft.named_parameters() if p.requires_grad)
    assert trainable and all("lora_" in n for n in trainable)

    S = cfg["max_sequence_length"]
    flat = synthetic.make_batches(cfg, rows, 1, seed=8, len_median=20.0, len_clip=(3, 80))[0]
    # finetune data holds targets only at each user's held-out last item (Finetune/transformer.jl:52-133);
    # that also keeps every weighted row inside the reference's top-k selection (mask_topk * B rows), the
    # condition under which losses and gradients do not depend on torch.topk's tie order
    u = flat["userid"]
    last = np.r_[u[1:] != u[:-1], True] & (u > 0)
    for k in flat:
        if k.endswith((".label", ".weight", ".position")):
            flat[k] = np.where(last, flat[k], 0).astype(flat[k].dtype)
    for m in (0, 1):
        assert (flat[f"{m}.{metric}.weight"] > 0).sum() <= cfg["mask_topk"] * rows
    batch = {k: v.reshape(rows, S) for k, v in flat.items()}
    for k in batch:
        if k.endswith(".position"):
            batch[k] = batch[k].astype(np.int64)            # to_device, transformer.py:182-183
    tw = oc.default_task_weights(finetune_medium=1, finetune_metric=metric)

    blob = {"config_json": np.frombuffer(json.dumps(cfg_ft).encode(), dtype=np.uint8), "rows": np.int64(rows),
            "task_weights": np.array(tw)}
    ft.eval()
    l32, g32, dm = run(ft, batch, tw, False)
    l16, g16, _ = run(ft, batch, tw, True)
    blob["eval/losses"], blob["eval/losses_bf16"] = np.array(l32), np.array(l16)
    for n in g32:
        blob["eval/grad/" + n], blob["eval/grad_bf16/" + n] = g32[n].numpy(), g16[n].float().numpy()

    ft.train()
    seed = 4242
    lt, gt, _ = run(ft, batch, tw, False, seed=seed)
    lt16, gt16, _ = run(ft, batch, tw, True, seed=seed)
    torch.manual_seed(seed)
    masks = [(F.dropout(torch.ones(rows, 2 * S, cfg["embed_dim"]), 0.1, True), F.dropout(torch.ones(rows, 2 * S, cfg["embed_dim"]), 0.1, True))
             for _ in range(cfg["num_layers"])]
    p = {k: v.detach().clone() for k, v in ft.state_dict().items() if not k.startswith("watch_head.")}
    lo, go, _ = oc.step_with_grads({k: torch.as_tensor(v) for k, v in batch.items()}, p, cfg_ft, None, tw, drops=masks)
    assert np.allclose([float(x) for x in lo], lt, rtol=2e-5), (lo, lt)   # the masks ARE the reference's masks
    for n in gt:
        assert float((go[n] - gt[n]).abs().max()) <= 1e-4 * float(gt[n].abs().max() + 1e-12), n
    blob["train/losses"], blob["train/losses_bf16"] = np.array(lt), np.array(lt16)
    for n in gt:
        blob["train/grad/" + n], blob["train/grad_bf16/" + n] = gt[n].numpy(), gt16[n].float().numpy()
    for i, (mq, mv) in enumerate(masks):
        blob[f"train/keep/{i}/q"] = np.packbits((mq > 0).numpy().reshape(-1))
        blob[f"train/keep/{i}/v"] = np.packbits((mv > 0).numpy().reshape(-1))

    for k, v in batch.items():
        blob["batch/" + k] = v
    for k, v in dm.items():
        blob["masked/" + k] = v.numpy()
    for k, v in ckpt.items():
        if not k.startswith("watch_head."):
            blob["ckpt/" + k] = v.numpy()
    for k, v in ft.state_dict().items():
        if "lora_" in k:
            blob["lora/" + k] = v.numpy()
    blob["fused_embedding"] = ft.item_embedding.fused_embedding.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), out)
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; eval losses", l32, "train losses", lt)


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY - loads the *real* reference model for pinning.

Executes /root/reference/notebooks/Training/transformer.model.py with the same
`exec` idiom its callers use (transformer.py:33-34, register.py:11-12), in a
prepared namespace.  Only usable inside the build container: the GPU box has no
/root/reference, so nothing under tests -m gpu / smoke() / bench.py imports this.
It exists to (1) generate tests/golden/*.npz and (2) validate oracle/cloze_oracle.py.

CPU shims (names injected, reference file untouched):
  * create_block_mask defaults to device="cuda"  -> partial(device="cpu")
  * CPU flex_attention has no backward            -> dense-mask SDPA, same
    scale (1/sqrt(hd)), same visibility mask, GQA through enable_gqa.
"""
from __future__ import annotations

import functools
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.attention.flex_attention import and_masks, create_block_mask, create_mask

REFERENCE_MODEL = "/root/reference/notebooks/Training/transformer.model.py"


def available():
    return os.path.exists(REFERENCE_MODEL)


def _dense_flex_attention(q, k, v, block_mask=None, enable_gqa=False, kernel_options=None, **_):
    B, H, Q, _ = q.shape
    mask = create_mask(block_mask.mask_mod, B, None, Q, k.shape[2], device=q.device)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, enable_gqa=enable_gqa)


def load_namespace(device="cpu"):
    g = {
        "torch": torch, "nn": nn, "F": F, "np": np, "h5py": None,
        "and_masks": and_masks,
        "create_block_mask": functools.partial(create_block_mask, device=device),
        "flex_attention": _dense_flex_attention,
    }
    with open(REFERENCE_MODEL) as f:
        exec(f.read(), g)
    return g


def build_model(config, seed, metadata=None, device="cpu"):
    g = load_namespace(device)
    torch.manual_seed(seed)
    model = g["RecommenderModel"](config)
    if metadata is not None:  # load_pretrained_embeddings semantics, transformer.model.py:379-389
        w = model.item_embedding.metadata_embedding.embedding.weight
        with torch.no_grad():
            w.zero_()
            w[: metadata.shape[0]].copy_(torch.as_tensor(metadata))
    return model.to(device)


def torch_batch(batch, device="cpu"):
    return {k: torch.as_tensor(np.ascontiguousarray(v)).to(device) for k, v in batch.items()}


def run_step(model, batch, task_weights, mask_seed, autocast=False, evaluate=False, device="cpu"):
    """One forward+backward exactly as train_epoch does it (transformer.py:256-272)."""
    d = torch_batch(batch, device)
    model.zero_grad(set_to_none=True)
    torch.manual_seed(mask_seed)
    ctx = torch.autocast("cpu" if device == "cpu" else "cuda", dtype=torch.bfloat16, enabled=autocast)
    with ctx:
        losses = model(d, evaluate)
        if evaluate:
            return losses, d, None
        loss = sum(losses[i] * task_weights[i] for i in range(len(losses)))
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return [l.detach() for l in losses], d, grads

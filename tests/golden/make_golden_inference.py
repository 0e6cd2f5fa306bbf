"""Writes tests/golden/inference_tiny.npz by running the REAL reference module in inference
mode (forward="inference", transformer.model.py:531-538) on batches shaped like
notebooks/Finetune/embed.py:74-131 builds them (rope_input_pos, per-candidate token_mask_ids).
Build container only:  python tests/golden/make_golden_inference.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh  # noqa: E402
from recommendersystem_b200 import synthetic  # noqa: E402


def make_batch(cfg, task, rng, lens, n_rank):
    """embed.py:74-131 `predict`, on synthetic tokens."""
    max_user = cfg["max_sequence_length"]
    S = max_user if task == "retrieval" else 2 * max_user
    V0, V1 = cfg["vocab_sizes"]["0_matchedid"], cfg["vocab_sizes"]["1_matchedid"]
    U = len(lens)
    d = {k: np.zeros((U, S), dt) for k, dt in (("userid", np.int32), ("time", np.float64), ("rope_input_pos", np.int32),
                                               ("token_mask_ids", np.int32), ("gender", np.int32), ("source", np.int32),
                                               ("matchedid", np.int32), ("status", np.int32), ("rating", np.float32),
                                               ("progress", np.float32))}
    for u, n in enumerate(lens):
        n = min(n, max_user - 1)
        extra = 1 if task == "retrieval" else n_rank[u]
        ts = np.sort(rng.uniform(cfg["min_ts"], cfg["max_ts"], n + 1))
        for i in range(n + extra):
            hist = i < n
            d["userid"][u, i] = u + 1
            d["time"][u, i] = ts[i] if hist else ts[n]
            d["gender"][u, i] = u % 4
            d["source"][u, i] = (u + 1) % 4
            d["rope_input_pos"][u, i] = i if hist else n
            d["token_mask_ids"][u, i] = i if (task == "ranking" and not hist) else 0
            if hist or task == "ranking":
                d["matchedid"][u, i] = rng.integers(0, V0 + V1)
            else:
                d["matchedid"][u, i] = -1                       # make_item(): masked item, embed.py:27-36
            d["status"][u, i] = rng.integers(0, 9) if hist else -1
            d["rating"][u, i] = float(rng.integers(0, 11)) if hist and rng.random() < 0.6 else 0.0
            d["progress"][u, i] = np.float32(rng.random()) if hist else 0.0
    return d


def main(out="inference_tiny.npz"):
    cfg, _ = synthetic.make_config("tiny", forward="inference")
    torch.set_num_threads(4)
    meta = synthetic.make_metadata(cfg, seed=3)
    model = rh.build_model(cfg, 4321, metadata=meta)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "periodic" in n or n.endswith("scale")):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        model.item_embedding.matchedid_embedding.embedding.weight[-1].normal_(0, 0.006, generator=g)
    model.eval()
    rng = np.random.default_rng(21)
    blob = {"config_json": np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)}
    for task, lens, n_rank in (("retrieval", [10, 37, 63, 1], None), ("ranking", [5, 40, 63], [3, 20, 64])):
        d = make_batch(cfg, task, rng, lens, n_rank)
        td = rh.torch_batch(d)
        with torch.no_grad():
            ref32 = model({k: v.clone() for k, v in td.items()}, task).float().numpy()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                ref16 = model({k: v.clone() for k, v in td.items()}, task).float().numpy()
        for k, v in d.items():
            blob[f"{task}/batch/{k}"] = v
        blob[f"{task}/out_fp32"], blob[f"{task}/out_bf16"] = ref32, ref16
        blob[f"{task}/lens"] = np.array(lens)
        print(task, ref32.shape, float(np.abs(ref32).mean()), float(np.abs(ref32 - ref16).max()))
    for k, v in model.state_dict().items():
        if not k.startswith("watch_head."):
            blob["param/" + k] = v.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), out)
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

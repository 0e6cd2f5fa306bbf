"""TEST INFRASTRUCTURE ONLY - CPU/PyTorch restatement of the reference cloze step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this file.  The product path (recommendersystem_b200)
never does, and fails loudly when libgraft.so is missing.

Pinned (tests/test_oracle_golden.py) against outputs of the reference's own
module run in the build container: tests/golden/*.npz, written by
tests/golden/make_golden.py through oracle/reference_harness.py.  The
reference holds no tests or golden vectors of its own.

Every function cites the reference lines it restates; paths are relative to
/root/reference/notebooks/Training/.  The restatement is functional (a params
dict with the reference's state_dict keys) instead of an nn.Module tree.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

MEDIUMS = (0, 1)            # transformer.model.py:1
METRICS = ("watch", "rating")  # transformer.model.py:2


# ---------------------------------------------------------------- masking ---
def mask_batch(d, randval, config):
    """transformer.model.py:417-462 (pre-train branch and finetune branch).

    `randval` replaces the in-module torch.rand draw (:437) so that callers
    control the seed.  Returns a new dict; inputs are not modified."""
    out = {k: v.clone() for k, v in d.items()}
    if config["finetune"]:
        pick = config["finetune_metric"]
        hit = torch.zeros_like(d["userid"], dtype=torch.bool)
        for k, v in d.items():
            if pick in k and k.endswith(".weight"):
                hit |= v > 0
        watch = hit if pick == "watch" else torch.zeros_like(hit)
        rating = hit if pick == "rating" else torch.zeros_like(hit)
    else:
        r = config["mask_rate"]
        watch = randval < r
        rating = (randval >= r) & (randval < 2 * r)
    either = watch | rating
    out["token_mask_ids"] = d["token_mask_ids"] * rating.to(d["token_mask_ids"].dtype)
    for k in d:
        tail = k.rsplit(".", 1)[-1] if "." in k else ""
        if tail in ("position", "label", "weight"):
            if "watch" in k:
                out[k] = torch.where(watch, d[k], torch.zeros_like(d[k]))
            elif "rating" in k:
                out[k] = torch.where(rating, d[k], torch.zeros_like(d[k]))
    out["matchedid"] = torch.where(watch, torch.full_like(d["matchedid"], -1), d["matchedid"])
    out["status"] = torch.where(either, torch.full_like(d["status"], -1), d["status"])
    out["rating"] = torch.where(either, torch.zeros_like(d["rating"]), d["rating"])
    out["progress"] = torch.where(either, torch.zeros_like(d["progress"]), d["progress"])
    return out


# ------------------------------------------------------------- embeddings ---
def _masked_lookup(table, idx):
    """transformer.model.py:15-24: -1 addresses the last row."""
    last = table.shape[0] - 1
    return F.embedding(torch.where(idx == -1, torch.full_like(idx, last), idx), table)


def action_features(d, p, config):
    """transformer.model.py:51-93: the 32 action features before the Linear.
    Time arithmetic stays in float64 until the cast (:55-66)."""
    lo, hi = config["min_ts"], config["max_ts"]
    ts = d["time"].clip(lo)
    ang = torch.stack([2 * np.pi * ts / 86400, 2 * np.pi * ts / (86400 * 7)], dim=-1).to(torch.float32)
    lin = ((ts - lo) / (hi - lo)).to(torch.float32).unsqueeze(-1)
    has = (d["rating"] != 0).int().unsqueeze(-1)
    z = has * ((d["rating"].unsqueeze(-1) - config["rating_mean"]) / config["rating_std"])
    return torch.cat(
        (
            lin,
            torch.cos(ang + p["action_embedding.periodic_time_cos"]),
            torch.sin(ang + p["action_embedding.periodic_time_sin"]),
            _masked_lookup(p["action_embedding.gender_embedding.embedding.weight"], d["gender"]),
            _masked_lookup(p["action_embedding.source_embedding.embedding.weight"], d["source"]),
            has,
            z,
            _masked_lookup(p["action_embedding.status_embedding.embedding.weight"], d["status"]),
            d["progress"].unsqueeze(-1),
        ),
        dim=-1,
    )


def item_table(p, lo=None, hi=None):
    """transformer.model.py:120-128 / 165-169: id rows + projected metadata rows; a fused
    table (finetune: ItemEmbedding.fuse() at load time, in fp32) is used as is."""
    sl = slice(lo, hi)
    if "item_embedding.fused_embedding" in p:
        return p["item_embedding.fused_embedding"][sl]
    return p["item_embedding.matchedid_embedding.embedding.weight"][sl] + F.linear(
        p["item_embedding.metadata_embedding.embedding.weight"][sl],
        p["item_embedding.projection_layer.weight"],
        p["item_embedding.projection_layer.bias"],
    )


def embed_tokens(d, p, config):
    """transformer.model.py:464-467 + 139-145 + 403-415: [B,2S,d], item token
    then action token per position."""
    act = F.linear(action_features(d, p, config), p["action_embedding.linear.weight"], p["action_embedding.linear.bias"])
    ids = d["matchedid"]
    if "item_embedding.fused_embedding" in p:  # transformer.model.py:139-142
        item = _masked_lookup(p["item_embedding.fused_embedding"], ids)
        B, S, D = item.shape
        return torch.stack((item, act), dim=2).reshape(B, 2 * S, D)
    item = _masked_lookup(p["item_embedding.matchedid_embedding.embedding.weight"], ids) + F.linear(
        _masked_lookup(p["item_embedding.metadata_embedding.embedding.weight"], ids),
        p["item_embedding.projection_layer.weight"],
        p["item_embedding.projection_layer.bias"],
    )
    B, S, D = item.shape
    return torch.stack((item, act), dim=2).reshape(B, 2 * S, D)


def visibility(userid, token_mask_ids):
    """transformer.model.py:479-487 evaluated densely: [B,2S,2S] bool,
    rows = queries.  Same user run and (key unmasked or same mask id)."""
    u = userid.repeat_interleave(2, dim=1)
    t = token_mask_ids.repeat_interleave(2, dim=1)
    same_user = u[:, :, None] == u[:, None, :]
    tok = (t[:, None, :] == 0) | (t[:, :, None] == t[:, None, :])
    return same_user & tok


# ------------------------------------------------------------------ trunk ---
def rms_norm(x, scale, eps=1e-5):
    """transformer.model.py:193-202."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return y.type_as(x) * scale


def rope_tables(head_dim, n, theta=500000.0):
    """transformer.model.py:173-179."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n, dtype=torch.float32), inv)
    return torch.cos(ang), torch.sin(ang)


def rope(x, cos, sin):
    """transformer.model.py:182-190; x [B,T,H,hd], pair-interleaved rotation.  cos/sin are
    [T, hd/2] (row positions) or [B, T, hd/2] (gathered by rope_input_pos, :336-339)."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xf[..., 0], xf[..., 1]
    if cos.dim() == 2:
        cos, sin = cos[None], sin[None]
    c, s = cos[:, :, None, :].float(), sin[:, :, None, :].float()
    return torch.stack((a * c - b * s, a * s + b * c), dim=-1).flatten(3).type_as(x)


def attention(x, p, pre, config, vis, cos, sin, drop=None, run_len=None):
    """transformer.model.py:256-286: softmax attention, scale 1/sqrt(hd); in finetune mode
    LoRA r=8, alpha=16 on q and v (:235-254, 263-271).  `drop` = (mask_q, mask_v) multiplies
    the LoRA input like nn.Dropout(0.1) in training (masks already scaled by 1/0.9); None = eval."""
    B, T, D = x.shape
    H, G = config["num_heads"], config["num_kv_heads"]
    hd = D // H
    q = F.linear(x, p[pre + "q_proj.weight"])
    k = F.linear(x, p[pre + "k_proj.weight"]).view(B, T, G, hd)
    v = F.linear(x, p[pre + "v_proj.weight"])
    if config.get("finetune"):
        xq, xv = (x, x) if drop is None else (x * drop[0], x * drop[1])
        q = q + F.linear(F.linear(xq, p[pre + "q_proj_lora_A.weight"]), p[pre + "q_proj_lora_B.weight"]) * 2.0
        v = v + F.linear(F.linear(xv, p[pre + "v_proj_lora_A.weight"]), p[pre + "v_proj_lora_B.weight"]) * 2.0
    q, v = q.view(B, T, H, hd), v.view(B, T, G, hd)
    q, k = rope(q, cos, sin), rope(k, cos, sin)
    if config.get("attention", "softmax") == "silu":
        o = silu_pointwise_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), vis, run_len)
    else:
        o = F.scaled_dot_product_attention(
            q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=vis[:, None], enable_gqa=H != G
        )
    return F.linear(o.transpose(1, 2).reshape(B, T, H * hd), p[pre + "output_proj.weight"])


def run_lengths(userid):
    """Tokens in each token's user run, [B, 2S] (item + action token per position)."""
    u = userid.repeat_interleave(2, dim=1)
    return (u[:, :, None] == u[:, None, :]).sum(-1)


def silu_pointwise_attention(q, k, v, vis, run_len):
    """BASELINE configs[2] variant, NOT in the reference (README.md:6 only cites HSTU, arXiv 2402.17152):
    point-wise attention out_t = sum_s silu(q_t.k_s / sqrt(hd)) / n_t * v_s over the reference's visibility
    (same user run, token-mask rule), n_t = number of tokens of t's user run (HSTU's 1/n), no softmax.
    q [B,H,T,hd], k/v [B,G,T,hd]; scores and silu in fp32, the weights are cast to q's dtype before the
    value product (as the probabilities of a fused kernel are).  Reference-unpinned: this function is the
    definition the CUDA kernels are held to."""
    B, H, T, hd = q.shape
    G = k.shape[1]
    k, v = k.repeat_interleave(H // G, 1), v.repeat_interleave(H // G, 1)
    s = (q.float() @ k.float().transpose(-1, -2)) / math.sqrt(hd)
    n = run_len.clamp(min=1).to(torch.float32)  # [B,T]
    a = torch.where(vis[:, None], F.silu(s), torch.zeros_like(s)) / n[:, None, :, None]
    return (a.to(q.dtype) @ v).to(q.dtype)


def swiglu(x, p, pre):
    """transformer.model.py:205-213."""
    return F.linear(F.silu(F.linear(x, p[pre + "w1.weight"])) * F.linear(x, p[pre + "w3.weight"]), p[pre + "w2.weight"])


def trunk(x, p, config, vis, positions=None, taps=None, drops=None, run_len=None):
    """transformer.model.py:297-343: pre-norm residual blocks + final norm.
    RoPE positions are token indices inside the row unless given (:336-339)."""
    T = x.shape[1]
    hd = config["embed_dim"] // config["num_heads"]
    cos, sin = rope_tables(hd, 2 * config["max_sequence_length"])
    cos, sin = cos.to(x.device), sin.to(x.device)
    if positions is None:
        cos, sin = cos[:T], sin[:T]
    else:
        cos, sin = cos[positions.long()], sin[positions.long()]  # [B, T, hd/2]
    for i in range(config["num_layers"]):
        pre = f"transformers.layers.{i}."
        x = x + attention(rms_norm(x, p[pre + "sa_norm.scale"]), p, pre + "attn.", config, vis, cos, sin,
                          drop=None if drops is None else drops[i], run_len=run_len)
        x = x + swiglu(rms_norm(x, p[pre + "mlp_norm.scale"]), p, pre + "mlp.")
        if taps is not None:
            taps[f"layer{i}"] = x
    return rms_norm(x, p["transformers.norm.scale"])


# ------------------------------------------------------------------ heads ---
def cloze_losses(hidden, d, p, config, evaluate=False):
    """transformer.model.py:499-529: [0.watch, 0.rating, 1.watch, 1.rating]."""
    B = d["userid"].shape[0]
    k_sel = config["mask_topk"] * B
    V0 = config["vocab_sizes"]["0_matchedid"]
    V = V0 + config["vocab_sizes"]["1_matchedid"]
    D = hidden.shape[-1]
    item_tok = hidden[:, 0::2].reshape(-1, D)
    act_tok = hidden[:, 1::2].reshape(-1, D)
    mu = config["rating_mean"]
    out = []
    for m in MEDIUMS:
        for metric in METRICS:
            w = d[f"{m}.{metric}.weight"].reshape(-1)
            _, at = torch.topk(w, k=k_sel)
            lab, pos, ws = d[f"{m}.{metric}.label"].reshape(-1)[at], d[f"{m}.{metric}.position"].reshape(-1)[at], w[at]
            den = ws.sum().clamp(min=1e-8)
            if metric == "watch":
                lo, hi = (0, V0) if m == 0 else (V0, V)
                logits = F.linear(item_tok[at], item_table(p, lo, hi))
                ce = F.cross_entropy(logits, pos.long(), reduction="none")
                out.append((ce * lab * ws).sum() / den)
            else:
                h = F.gelu(F.linear(act_tok[at], p["rating_head.0.weight"], p["rating_head.0.bias"]))
                pred = F.linear(h, p["rating_head.2.weight"], p["rating_head.2.bias"]).reshape(-1)
                y = lab - mu
                if evaluate:  # three moments, transformer.model.py:395-401
                    out.append([(torch.square(c * pred - y) * ws).sum() / den for c in (1, 0, -1)])
                else:
                    out.append((torch.square(pred - y) * ws).sum() / den)
    return out


def step(d, p, config, randval, evaluate=False, taps=None, drops=None):
    """transformer.model.py:493-529 train_forward.  d: dict of flat or [B,S]
    tensors; returns (losses, masked batch)."""
    S = config["max_sequence_length"]
    d = {k: v.reshape(-1, S) for k, v in d.items()}
    dm = mask_batch(d, randval.reshape(-1, S) if randval is not None else None, config)
    x0 = embed_tokens(dm, p, config)
    vis = visibility(dm["userid"], dm["token_mask_ids"])
    if taps is not None:
        taps["x0"] = x0
    hidden = trunk(x0, p, config, vis, taps=taps, drops=drops, run_len=run_lengths(dm["userid"]))
    if taps is not None:
        taps["hidden"] = hidden
    return cloze_losses(hidden, dm, p, config, evaluate), dm


def inference_forward(d, p, config, task):
    """transformer.model.py:531-538 + 464-491: d holds [U, S'] arrays; `rope_input_pos`
    (optional) gives the RoPE position of each history position (tokens 2p, 2p+1)."""
    x0 = embed_tokens(d, p, config)
    vis = visibility(d["userid"], d["token_mask_ids"])
    pos = None
    if "rope_input_pos" in d:
        rp = d["rope_input_pos"]
        pos = torch.stack((2 * rp, 2 * rp + 1), dim=2).reshape(rp.shape[0], -1)
    e = trunk(x0, p, config, vis, positions=pos, run_len=run_lengths(d["userid"]))
    if task == "retrieval":
        return e
    assert task == "ranking"
    h = F.gelu(F.linear(e, p["rating_head.0.weight"], p["rating_head.0.bias"]))
    return F.linear(h, p["rating_head.2.weight"], p["rating_head.2.bias"])


def step_with_grads(d, p, config, randval, task_weights, autocast=False, drops=None):
    """Forward + backward as transformer.py:256-272 drives it; returns
    (losses, grads dict keyed like p for float leaves that require grad)."""
    leaves = {}
    for k, v in p.items():
        t = v.detach().clone()
        trainable = t.is_floating_point() and k != "item_embedding.metadata_embedding.embedding.weight"
        if config.get("finetune"):  # transformer.model.py:361-371: only the LoRA matrices train
            trainable = "lora_" in k
        if trainable:
            t.requires_grad_(True)
        leaves[k] = t
    dev = "cuda" if next(iter(leaves.values())).is_cuda else "cpu"
    with torch.autocast(dev, dtype=torch.bfloat16, enabled=autocast):
        losses, dm = step(d, leaves, config, randval, drops=drops)
        total = sum(l * w for l, w in zip(losses, task_weights))
    total.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.requires_grad and v.grad is not None}
    return [l.detach() for l in losses], grads, dm


def default_task_weights(finetune_medium=None, finetune_metric=None):
    """transformer.py:379-406 `make_task_weights`: order [0.watch, 0.rating, 1.watch, 1.rating]."""
    scale = [4.618602403897067, 1.1958987168236102, 2.5443243303769867, 1.0527565486045412]
    metric_w = {"watch": 1, "rating": 0.25} if finetune_metric is None else {m: float(m == finetune_metric) for m in METRICS}
    medium_w = {0: 0.25, 1: 1} if finetune_medium is None else {finetune_medium: 1, 1 - finetune_medium: 0}
    w = [medium_w[m] * metric_w[k] for m in MEDIUMS for k in METRICS]
    tot = sum(w)
    return [(x / tot) / s for x, s in zip(w, scale)]


# ---------------------------------------------------------------- packing ---
def jagged_offsets(userid):
    """transformer.py:53-68 `get_index_permutation`: run boundaries of equal
    userid.  Returns int64 offsets [n_runs+1] with offsets[0]=0, [-1]=len."""
    u = np.asarray(userid)
    if u.size == 0:
        return np.zeros(1, dtype=np.int64)
    change = np.where(u[:-1] != u[1:])[0] + 1
    return np.concatenate(([0], change, [u.size])).astype(np.int64)


def block_permutation(offsets, block_order):
    """transformer.py:58-67: index permutation that moves whole user blocks."""
    parts = [np.arange(offsets[b], offsets[b + 1]) for b in block_order]
    return np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, dtype=np.int64)


def concat_pad(per_user, multiple):
    """transformer.jl:144-163 `concat`: back-to-back, zero pad to a multiple."""
    keys = list(per_user[0].keys())
    n = sum(len(u["userid"]) for u in per_user)
    total = n + (-n) % multiple
    out = {}
    for k in keys:
        buf = np.zeros(total, dtype=per_user[0][k].dtype)
        buf[:n] = np.concatenate([u[k] for u in per_user])
        out[k] = buf
    return out

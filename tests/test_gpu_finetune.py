"""Finetune mode on the GPU against the reference's own finetune-mode
outputs (tests/golden/finetune_tiny.npz, tests/golden/make_golden_finetune.py): LoRA r=8 on q,v,
frozen fused item table, deterministic finetune masking, LoRA-input dropout fed the reference's masks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_oracle_golden import _finetune_golden  # noqa: E402

DEV = "cuda:0"


def _model(z, cfg):
    from recommendersystem_b200.model import RecommenderModel
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    ckpt = {k[5:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("ckpt/")}
    for k in [k for k in ckpt if k.startswith("item_embedding.")]:
        ckpt["watch_head." + k] = ckpt[k]
    assert not model.item_embedding.fused
    model.load_state_dict(ckpt, strict=False)       # transformer.py:650-656: the post-hook fuses the item table
    assert model.item_embedding.fused
    lora = {k[5:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("lora/")}
    missing, unexpected = model.load_state_dict(lora, strict=False)
    assert not unexpected
    return model.to(DEV)


def test_finetune_surface():
    z, cfg, p, batch, masks = _finetune_golden()
    model = _model(z, cfg)
    trainable = sorted(n for n, q in model.named_parameters() if q.requires_grad)
    assert trainable == sorted(k[5:] for k in z.files if k.startswith("lora/"))
    sd = model.state_dict()
    assert "item_embedding.fused_embedding" in sd and "watch_head.item_embedding.fused_embedding" in sd
    assert not any("matchedid_embedding" in k or "projection_layer" in k for k in sd)
    assert torch.equal(sd["item_embedding.fused_embedding"].cpu(), torch.as_tensor(z["fused_embedding"]))
    with pytest.raises(AssertionError):             # a fused checkpoint must not be loaded into an un-fused model
        from recommendersystem_b200.model import RecommenderModel
        RecommenderModel(cfg).load_state_dict({"item_embedding.fused_embedding": torch.zeros(1)}, strict=False)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_finetune_step_matches_reference(mode):
    z, cfg, p, batch, masks = _finetune_golden()
    model = _model(z, cfg)
    model.train(mode == "train")
    if mode == "train":
        model.lora_mask = lambda layer, which, shape, dev: masks[layer][0 if which == "q" else 1].to(dev)
    d = {k: v.to(DEV) for k, v in batch.items()}
    losses = model(d, False)
    tw = [float(x) for x in z["task_weights"]]
    sum(l * w for l, w in zip(losses, tw)).backward()
    got = np.array([float(l) for l in losses])
    ref16, ref32 = z[f"{mode}/losses_bf16"], z[f"{mode}/losses"]
    print(mode, "losses", got, "ref bf16", ref16, "ref fp32", ref32)
    assert np.allclose(got, ref16, rtol=1e-2, atol=1e-6) and np.allclose(got, ref32, rtol=1e-2, atol=1e-6)
    for k in d:                                      # finetune masking edits the batch like the reference
        assert np.array_equal(d[k].cpu().numpy().reshape(-1), z["masked/" + k].reshape(-1)), k
    worst, bad = 0.0, []
    for n, q in model.named_parameters():
        if not q.requires_grad:
            assert q.grad is None, n
            continue
        ref = z[f"{mode}/grad/{n}"]
        g = q.grad.float().cpu().numpy()
        rel = np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-20)
        gap = np.linalg.norm(z[f"{mode}/grad_bf16/{n}"] - ref) / max(np.linalg.norm(ref), 1e-20)
        worst = max(worst, rel)
        print(f"{mode} grad rel {rel:.3e} (reference bf16 gap {gap:.3e}) |g| {np.linalg.norm(g):.3e} |ref| {np.linalg.norm(ref):.3e} {n}")
        if not rel < max(2e-2, 2 * gap):                 # bf16 tolerance; never worse than 2x the reference's own bf16 gap
            bad.append((n, rel, gap))
    print(mode, "worst LoRA grad rel-L2 vs fp32 reference", worst)
    assert not bad, bad


def test_finetuned_inference_merges_lora():
    """inference with finetune=True (embed.py's production setting): fused table buffer at construction,
    LoRA merged into the projection weights; compared with the oracle in fp32."""
    from oracle import cloze_oracle as oc
    from recommendersystem_b200.model import RecommenderModel
    z, cfg, p, batch, masks = _finetune_golden()
    cfg_inf = dict(cfg, forward="inference")
    model = RecommenderModel(cfg_inf)
    assert model.item_embedding.fused
    sd = {k: v for k, v in p.items()}
    sd["watch_head.item_embedding.fused_embedding"] = sd["item_embedding.fused_embedding"]
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    d = {k: batch[k] for k in ("userid", "time", "token_mask_ids", "gender", "source", "matchedid", "status", "rating", "progress")}
    with torch.no_grad():
        ref = oc.inference_forward(d, p, cfg_inf, "retrieval").numpy()
        out = model({k: v.to(DEV) for k, v in d.items()}, "retrieval").float().cpu().numpy()
    u = np.repeat(d["userid"].numpy(), 2, axis=1) > 0
    rel = np.linalg.norm(out[u] - ref[u]) / np.linalg.norm(ref[u])
    print("finetuned inference rel vs fp32 oracle", rel)
    assert rel < 1e-2

"""Inference forward (transformer.model.py:531-538) against outputs of the
reference's own module (tests/golden/inference_tiny.npz, tests/golden/make_golden_inference.py):
batches shaped as notebooks/Finetune/embed.py:74-131 builds them - rope_input_pos, a masked
query item for retrieval, per-candidate token_mask_ids for ranking."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inference_tiny.npz"))
    cfg = json.loads(bytes(z["config_json"]).decode())
    params = {k[6:]: z[k] for k in z.files if k.startswith("param/")}
    return z, cfg, params


def _batch(z, task):
    pre = f"{task}/batch/"
    return {k[len(pre):]: torch.as_tensor(z[k]).to(DEV) for k in z.files if k.startswith(pre)}


@pytest.mark.parametrize("task", ["retrieval", "ranking"])
def test_inference_matches_reference(task):
    from recommendersystem_b200.testing import model_from_params
    z, cfg, params = _load()
    model = model_from_params(cfg, params, DEV).eval()
    d = _batch(z, task)
    keep = {k: v.clone() for k, v in d.items()}
    with torch.no_grad():
        out = model(d, task).float().cpu().numpy()
    ref16, ref32 = z[f"{task}/out_bf16"], z[f"{task}/out_fp32"]
    assert out.shape == ref32.shape
    for k in d:  # inference does not edit the batch
        assert torch.equal(d[k], keep[k]), k
    # rows that belong to users (padding rows of userid 0 form one garbage segment in both)
    u = np.repeat(z[f"{task}/batch/userid"], 2, axis=1) > 0
    got, r16, r32 = out[u], ref16[u], ref32[u]
    rel32 = np.linalg.norm(got - r32) / np.linalg.norm(r32)
    rel16 = np.linalg.norm(got - r16) / np.linalg.norm(r16)
    ref_gap = np.linalg.norm(r16 - r32) / np.linalg.norm(r32)
    print(f"{task}: rel vs fp32 ref {rel32:.3e}, vs bf16-autocast ref {rel16:.3e} (bf16 ref vs fp32 ref {ref_gap:.3e})")
    assert rel32 < 1e-2 and rel16 < 1e-2   # bf16 tolerance of the north star
    if task == "retrieval":  # the row the server reads: token 2N of each user (embed.py:143-150)
        for i, n in enumerate(z["retrieval/lens"]):
            a, b = out[i, 2 * int(n)], ref32[i, 2 * int(n)]
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-2


def test_fused_item_table_is_fuse():
    """ItemEmbedding.fuse() (transformer.model.py:120-133), what register.py:14-38 exports."""
    from recommendersystem_b200.testing import model_from_params
    z, cfg, params = _load()
    model = model_from_params(cfg, params, DEV)
    t = model.fused_item_table().cpu()
    ref = torch.as_tensor(params["item_embedding.matchedid_embedding.embedding.weight"]) + torch.nn.functional.linear(
        torch.as_tensor(params["item_embedding.metadata_embedding.embedding.weight"]),
        torch.as_tensor(params["item_embedding.projection_layer.weight"]), torch.as_tensor(params["item_embedding.projection_layer.bias"]))
    assert float((t - ref).norm() / ref.norm()) < 3e-3
    assert torch.equal(t[-1], ref[-1].float()) or float((t[-1] - ref[-1]).abs().max()) < 1e-3  # mask row: id row + bias


def test_retrieval_scores_match_softmax_of_catalog_logits():
    """Finetune/embed.jl:86-90 compute_retrieval: softmax(W u)[idxs] with W = the exported item table slice."""
    from recommendersystem_b200.testing import model_from_params
    z, cfg, params = _load()
    model = model_from_params(cfg, params, DEV).eval()
    g = torch.Generator().manual_seed(4)
    u = (torch.randn(5, cfg["embed_dim"], generator=g) * 2.0).to(DEV)
    table = model.fused_item_table()
    V0 = cfg["vocab_sizes"]["0_matchedid"]
    for medium, (lo, hi) in enumerate(((0, V0), (V0, table.shape[0] - 1))):
        want = torch.softmax(u @ table[lo:hi].t(), dim=-1)
        got = model.retrieval_scores(u, medium)
        assert got.shape == want.shape
        assert float((got - want).norm() / want.norm()) < 1e-2
        assert float((got.sum(-1) - 1).abs().max()) < 1e-2
        idxs = torch.tensor([3, 0, hi - lo - 1, 7])
        sel = model.retrieval_scores(u, medium, idxs)
        assert float((sel - want[:, idxs.to(DEV)]).norm() / want[:, idxs.to(DEV)].norm()) < 1e-2

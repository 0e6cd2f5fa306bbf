"""The drop-in at its boundary, driven the way the reference's own callers drive it: the model file is exec'd into the
caller's namespace (Training/transformer.py:33-34, Finetune/register.py:11-12, Finetune/embed.py:164-165), a checkpoint is
loaded, and modules are called directly - register.py:29 exports the catalog with `model.item_embedding(torch.arange(...))`."""
import io
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _exec_model_file():
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    with open(os.path.join(ROOT, "integration", "transformer.model.py")) as f:
        exec(f.read(), ns)
    return ns


def _golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inference_tiny.npz"))
    cfg = json.loads(bytes(z["config_json"]).decode())
    params = {k[6:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("param/")}
    for k in list(params):
        if k.startswith("item_embedding."):
            params["watch_head." + k] = params[k]
    return cfg, params


def test_register_py_flow():
    """Finetune/register.py:14-38 line by line, with device = cuda (the drop-in has no CPU path)."""
    ns = _exec_model_file()
    cfg, params = _golden()
    buf = io.BytesIO()
    torch.save({"config": dict(cfg, forward="train"), "model": params}, buf)
    buf.seek(0)
    checkpoint = torch.load(buf, weights_only=False, map_location=DEV)
    config = checkpoint["config"]
    config["forward"] = "inference"
    model = ns["RecommenderModel"](config)
    model.load_state_dict(checkpoint["model"])
    model = model.to(DEV)
    model.eval()
    n_0 = model.config["vocab_sizes"]["0_matchedid"]
    n_1 = model.config["vocab_sizes"]["1_matchedid"]
    embs = model.item_embedding(torch.arange(0, n_0 + n_1))
    ret = {"0.watch.weight": embs[:n_0, :].detach().cpu().numpy(), "1.watch.weight": embs[n_0:(n_0 + n_1), :].detach().cpu().numpy()}
    ref = params["item_embedding.matchedid_embedding.embedding.weight"] + F.linear(
        params["item_embedding.metadata_embedding.embedding.weight"], params["item_embedding.projection_layer.weight"],
        params["item_embedding.projection_layer.bias"])
    got = np.concatenate([ret["0.watch.weight"], ret["1.watch.weight"]])
    assert got.shape == (n_0 + n_1, config["embed_dim"])
    err = np.linalg.norm(got - ref[:-1].numpy()) / np.linalg.norm(ref[:-1].numpy())
    assert err < 3e-3, err   # the projection runs in bf16, as the autocast reference's does
    # -1 reads the mask row (ref:23-24, 141)
    m = model.item_embedding(torch.tensor([-1, 0], device=DEV))
    assert torch.equal(m[0], model.fused_item_table()[-1])


def test_module_forwards_match_reference_formulas():
    ns = _exec_model_file()
    cfg, params = _golden()
    model = ns["RecommenderModel"](dict(cfg, forward="inference"))
    model.load_state_dict(params)
    model = model.to(DEV).eval()
    # MaskedEmbedding (ref:15-24): bit-exact gather, any integer dtype
    se = model.action_embedding.status_embedding
    idx = torch.tensor([[0, 3, -1], [8, 1, 2]], device=DEV)
    ref = F.embedding(torch.where(idx == -1, se.vocab_size - 1, idx), se.embedding.weight)
    assert torch.equal(se(idx), ref) and torch.equal(se(idx.int()), ref)
    # DualItemEmbedding (ref:153-170): logits against one medium's slice of the tied table
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(37, cfg["embed_dim"], generator=g) * 0.5).to(DEV)
    table = model.fused_item_table()
    V0 = cfg["vocab_sizes"]["0_matchedid"]
    for medium, (lo, hi) in enumerate(((0, V0), (V0, table.shape[0] - 1))):
        logits = model.watch_head((x, medium))
        assert logits.shape == (37, hi - lo) and logits.dtype == torch.bfloat16
        want = x.bfloat16().float() @ table[lo:hi].bfloat16().float().t()
        assert float((logits.float() - want).norm() / want.norm()) < 1e-2
    # ActionEmbedding (ref:51-94) against the oracle's restatement
    from oracle import cloze_oracle as oc
    z = np.load(os.path.join(ROOT, "tests", "golden", "inference_tiny.npz"))
    d = {k[len("retrieval/batch/"):]: torch.as_tensor(z[k]).to(DEV) for k in z.files if k.startswith("retrieval/batch/")}
    got = model.action_embedding(d)
    p = {k: v.to(DEV) for k, v in params.items()}
    want = F.linear(oc.action_features(d, p, cfg), p["action_embedding.linear.weight"], p["action_embedding.linear.bias"])
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    assert float((got.float() - want).norm() / want.norm()) < 1e-2


def test_dtype_normalisation_and_flags():
    """int64 ids / float32 time are converted (they used to be reinterpreted); the status word reports a recurring userid
    and more weighted rows than mask_topk * B instead of silently changing the result."""
    from conftest import load_golden
    from recommendersystem_b200 import lib as L
    from recommendersystem_b200.testing import batch_to_device, model_from_params
    g = load_golden()
    model = model_from_params(g["config"], g["param"], DEV)
    rv = torch.as_tensor(g["randval"]).to(DEV)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    base = [float(x) for x in model(batch_to_device(g["batch"], DEV), False)]
    d = batch_to_device(g["batch"], DEV)
    for k in d:
        if d[k].dtype == torch.int32:
            d[k] = d[k].long()
    got = [float(x) for x in model(d, False)]
    assert got == base
    L.status_flags(clear=True)
    model(batch_to_device(g["batch"], DEV), False)
    assert L.status_flags() == 0
    # a userid that comes back later in the same row
    d = batch_to_device(g["batch"], DEV)
    S = g["config"]["max_sequence_length"]
    u = d["userid"].view(-1, S)
    u[0, -1] = u[0, 0]
    if int(u[0, -2]) == int(u[0, 0]):
        u[0, -2] = 12345
    model(d, False)
    assert L.status_flags() & L.FLAG_USERID_RECURS
    # more weighted rows than the selection holds
    d = batch_to_device(g["batch"], DEV)
    d["0.watch.weight"].fill_(1.0)
    model.mask_rand = lambda shape, dev: torch.zeros(shape, device=dev)   # every position is a watch target
    model(d, False)
    assert L.status_flags() & L.FLAG_SELECT_OVERFLOW
    d = batch_to_device(g["batch"], DEV); d["0.watch.weight"].fill_(1.0); model(d, False)
    with pytest.raises(L.GraftError):
        L.check_status()
    assert L.status_flags() == 0   # check_status cleared the word

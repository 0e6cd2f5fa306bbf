"""Pins oracle/cloze_oracle.py to outputs of the reference's own module
(tests/golden/cloze_tiny.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import cloze_oracle as oc


def _tensors(d):
    return {k: torch.as_tensor(np.ascontiguousarray(v)) for k, v in d.items()}


def test_mask_matches_reference(golden):
    cfg = golden["config"]
    S = cfg["max_sequence_length"]
    d = {k: v.reshape(-1, S) for k, v in _tensors(golden["batch"]).items()}
    dm = oc.mask_batch(d, torch.as_tensor(golden["randval"]), cfg)
    for k, ref in golden["masked"].items():
        got = dm[k].reshape(-1).numpy()
        assert got.dtype == ref.dtype, k
        assert np.array_equal(got, ref), k


def test_losses_and_grads_fp32(golden):
    cfg = golden["config"]
    p = _tensors(golden["param"])
    losses, grads, _ = oc.step_with_grads(_tensors(golden["batch"]), p, cfg, torch.as_tensor(golden["randval"]),
                                          golden["task_weights"])
    np.testing.assert_allclose([float(x) for x in losses], golden["losses"], rtol=2e-5)
    assert set(grads) == set(golden["grad"])
    for k, ref in golden["grad"].items():
        got = grads[k].numpy()
        den = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got - ref).max() / den < 1e-4, (k, np.abs(got - ref).max() / den)


def test_losses_bf16_autocast(golden):
    cfg = golden["config"]
    p = _tensors(golden["param"])
    losses, grads, _ = oc.step_with_grads(_tensors(golden["batch"]), p, cfg, torch.as_tensor(golden["randval"]),
                                          golden["task_weights"], autocast=True)
    np.testing.assert_allclose([float(x) for x in losses], golden["losses_bf16"], rtol=2e-3)
    for k, ref in golden["grad_bf16"].items():
        got = grads[k].float().numpy()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12)
        assert rel < 2e-2, (k, rel)


def test_eval_moments(golden):
    cfg = golden["config"]
    p = _tensors(golden["param"])
    with torch.no_grad():
        out, _ = oc.step(_tensors(golden["batch"]), p, cfg, torch.as_tensor(golden["randval"]), evaluate=True)
    flat = []
    for e in out:
        flat += [float(x) for x in e] if isinstance(e, list) else [float(e)]
    np.testing.assert_allclose(flat, golden["eval_flat"], rtol=2e-5)


def test_jagged_offsets_and_block_permutation():
    u = np.array([3, 3, 3, 1, 1, 7, 0, 0, 0, 0], dtype=np.int32)
    off = oc.jagged_offsets(u)
    assert off.tolist() == [0, 3, 5, 6, 10]
    assert oc.jagged_offsets(np.zeros(0, np.int32)).tolist() == [0]
    assert oc.jagged_offsets(np.array([5], np.int32)).tolist() == [0, 1]
    perm = oc.block_permutation(off, [2, 0, 3, 1])
    assert perm.tolist() == [5, 0, 1, 2, 6, 7, 8, 9, 3, 4]
    # restates transformer.py:53-68 literally for a cross-check
    change = np.where(u[:-1] != u[1:])[0] + 1
    blocks = np.split(np.arange(len(u)), change)
    assert np.concatenate([blocks[b] for b in [2, 0, 3, 1]]).tolist() == perm.tolist()


def test_inference_forward_matches_reference():
    """oracle.inference_forward vs the reference module in inference mode (rope_input_pos,
    per-candidate token_mask_ids): tests/golden/inference_tiny.npz."""
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference_tiny.npz"))
    cfg = json.loads(bytes(z["config_json"]).decode())
    p = {k[6:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("param/")}
    for task in ("retrieval", "ranking"):
        pre = f"{task}/batch/"
        d = {k[len(pre):]: torch.as_tensor(z[k]) for k in z.files if k.startswith(pre)}
        with torch.no_grad():
            out = oc.inference_forward(d, p, cfg, task).numpy()
        ref = z[f"{task}/out_fp32"]
        assert np.abs(out - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), task


def _finetune_golden():
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "finetune_tiny.npz"))
    cfg = json.loads(bytes(z["config_json"]).decode())
    p = {k[5:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("ckpt/")}
    for k in [k for k in p if k.startswith("item_embedding.")]:
        del p[k]                                            # fused away at load time (transformer.model.py:120-133)
    p["item_embedding.fused_embedding"] = torch.as_tensor(z["fused_embedding"])
    p.update({k[5:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("lora/")})
    batch = {k[6:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("batch/")}
    S, dmod, rows = cfg["max_sequence_length"], cfg["embed_dim"], int(z["rows"])
    masks = []
    for i in range(cfg["num_layers"]):
        mk = [torch.as_tensor(np.unpackbits(z[f"train/keep/{i}/{w}"])[: rows * 2 * S * dmod].reshape(rows, 2 * S, dmod).astype(np.float32))
              / 0.9 for w in "qv"]
        masks.append(tuple(mk))
    return z, cfg, p, batch, masks


def test_finetune_mode_matches_reference():
    """LoRA on q,v, fused frozen table, finetune masking, with and without the LoRA dropout masks:
    oracle vs the reference module in finetune mode (tests/golden/finetune_tiny.npz)."""
    z, cfg, p, batch, masks = _finetune_golden()
    tw = [float(x) for x in z["task_weights"]]
    for mode, drops in (("eval", None), ("train", masks)):
        losses, grads, dm = oc.step_with_grads(batch, p, cfg, None, tw, drops=drops)
        np.testing.assert_allclose([float(x) for x in losses], z[f"{mode}/losses"], rtol=2e-5, atol=1e-7)
        names = [k[len(mode) + 6:] for k in z.files if k.startswith(f"{mode}/grad/")]
        assert sorted(names) == sorted(grads) and all("lora_" in n for n in names)
        for n in names:
            ref = z[f"{mode}/grad/{n}"]
            assert np.abs(grads[n].numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-12), (mode, n)
        if mode == "eval":
            for k in dm:
                assert np.array_equal(dm[k].numpy().reshape(-1), z["masked/" + k].reshape(-1)), k

"""Helpers shared by tests, smoke() and bench.py to drive the model on fixtures."""
import numpy as np
import torch

from .model import RecommenderModel


def model_from_params(config, params, device):
    """Builds the drop-in model and loads a reference-keyed state dict (numpy)."""
    model = RecommenderModel(config)
    sd = {k: torch.as_tensor(np.ascontiguousarray(v)) for k, v in params.items()}
    for k in list(sd):  # tied head aliases, transformer.model.py:354
        if k.startswith("item_embedding."):
            sd["watch_head." + k] = sd[k]
    model.load_state_dict(sd)
    return model.to(device)


def batch_to_device(batch, device):
    return {k: torch.as_tensor(np.ascontiguousarray(v)).to(device) for k, v in batch.items()}


def run_golden_step(g, device="cuda:0", evaluate=False):
    """One forward(+backward) on the golden fixture's inputs, mask draw injected."""
    model = model_from_params(g["config"], g["param"], device)
    rv = torch.as_tensor(g["randval"]).to(device)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    d = batch_to_device(g["batch"], device)
    if evaluate:
        with torch.no_grad():
            return model(d, True), None, d
    losses = model(d, False)
    total = sum(l * w for l, w in zip(losses, g["task_weights"]))
    total.backward()
    grads = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
    return [l.detach() for l in losses], grads, d

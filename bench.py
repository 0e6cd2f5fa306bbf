"""bench.py - cloze training step throughput (train interactions/s), BASELINE.json metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2]

One step = one micro-batch of the reference's pre-train step on one GPU:
H2D of the 27-array batch (e2e leg) -> mask -> embed -> 8-block stack -> cloze +
rating heads -> backward (incl. deterministic scatter-add) -> data-parallel
all-reduce of the flat gradient -> clip -> AdamW.  Nothing is skipped.
Workload at N=1: BASELINE configs[1] = C2 (64 rows x 1024 positions, d=768, 8
blocks, 100k catalog, bf16 with fp32 master weights/residual), synthetic
histories.  N>1: weak scaling, each rank its own 64 rows.

`--impl reference`: the reference's CPU path (the pinned oracle port of its
PyTorch module; the reference itself cannot be imported on the GPU box) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from recommendersystem_b200 import synthetic  # noqa: E402

METRIC = "train interactions/sec (cloze step)"
UNIT = "interactions/s"


def workload_config(name, world, graphed):
    """`config` of the JSON line: the same object for the GPU arm and the reference arm."""
    cfg, rows = synthetic.make_config(name)
    S = cfg["max_sequence_length"]
    extra = {"C3": "; HSTU-style point-wise SiLU attention, rating-only task weights", "C4": "; vocab-parallel catalog logits/CE",
             "C4L": "; vocab-parallel catalog logits/CE, 1M-item catalog (metadata width 512)",
             "C5": "; short histories (5-64 positions), Zipf 1.2 items"}.get(name, "")
    out = {"workload": f"{name}: {rows} rows x {S} positions/GPU, d={cfg['embed_dim']}, L={cfg['num_layers']}, "
                       f"H/G={cfg['num_heads']}/{cfg['num_kv_heads']}, FFN={cfg['intermediate_dim']}, catalog="
                       f"{cfg['vocab_sizes']['0_matchedid'] + cfg['vocab_sizes']['1_matchedid']}, metadata={cfg['metadata_emb_size']}; "
                       f"fwd+bwd+allreduce+clip+AdamW; synthetic lognormal histories, Zipf items{extra}",
           "global_batch_interactions": world * rows * S, "parallelism": f"dp{world}",
           "l2": "per-step working set (>25 GB of activations) far exceeds the 126 MB L2; 4 distinct batches rotate"}
    if graphed is not None:
        out["cuda_graph"] = "fwd+bwd replayed from one CUDA graph; optimiser tail launched eagerly" if graphed else "off"
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm": j["hbm_gbs"], "tensor": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "src": "fallback"}


def gemm_traffic():
    """DRAM bytes (read + write) of one launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/*gemm_traffic.json, written from the .ncu-rep by tools/ncu_summary.py); None if absent."""
    import glob
    fs = sorted(glob.glob(os.path.join(ROOT, "profiles", "*gemm_traffic.json")))
    if not fs:
        return None
    return json.load(open(fs[-1]))


# ------------------------------------------------------------------ CPU reference leg --
def cpu_reference(cfg_name, rows_sample, steps, warmup, threads=None, budget_s=None):
    """Oracle port of the reference step (fwd+bwd, bf16 autocast like the trainer) on the host cores.
    Bounded sample: `rows_sample` rows of the named config per step (every step is a full fwd+bwd of those rows,
    including the per-step [catalog x metadata] projection the reference's tied head performs).  With `budget_s`, the
    sample is halved (8 -> 4 -> 2 -> 1 rows) until warmup + steps fit the budget, judged from one probe step; the
    steps and warm-ups asked for are always run in full."""
    from oracle import cloze_oracle as oc
    from recommendersystem_b200.model import RecommenderModel
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg, _ = synthetic.make_config(cfg_name)
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=synthetic.make_metadata(cfg))
    p = {k: v.detach() for k, v in model.state_dict().items() if not k.startswith("watch_head.")}
    cfg_s = dict(cfg)
    tw = oc.default_task_weights(finetune_metric="rating") if cfg_name == "C3" else oc.default_task_weights()
    S = cfg["max_sequence_length"]
    kw = synthetic.sampler_kwargs(cfg_name)

    def run(rows, n_steps, n_warm):
        batches = synthetic.make_batches(cfg, rows, 2, seed=0, **kw)
        times = []
        for i in range(n_warm + n_steps):
            d = {k: torch.as_tensor(v) for k, v in batches[i % 2].items()}
            torch.manual_seed(i)
            rv = torch.rand(rows, S)
            t0 = time.perf_counter()
            oc.step_with_grads(d, p, cfg_s, rv, tw, autocast=True)
            dt = time.perf_counter() - t0
            if i >= n_warm:
                times.append(dt)
        return times

    rows = rows_sample
    if budget_s is not None:
        probe = run(rows, 1, 0)[0]
        while rows > 1 and probe * (steps + warmup) > budget_s:
            rows //= 2
            probe = run(rows, 1, 0)[0]
    times = run(rows, steps, warmup)
    n = rows * S
    t = float(np.median(times))
    return {"value": n / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rows} of the workload's rows ({n} interactions) per step, {steps} steps after {warmup} warm-up, "
                      f"bf16 autocast fwd+bwd incl. the per-step catalog projection (amortised over {rows} rows, not the workload's), median step",
            "ms_per_step": t * 1e3, "steps": steps, "warmup": warmup, "rows": rows}


# ------------------------------------------------------------------ clocks sampling ----
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ GPU leg ------------
def gpu_bench(args):
    from recommendersystem_b200 import lib as L
    from recommendersystem_b200.model import RecommenderModel
    from recommendersystem_b200.optim import FlatAdamW, make_task_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    cfg, rows = synthetic.make_config(args.config)
    torch.manual_seed(0)
    model = RecommenderModel(cfg)
    model.load_pretrained_embeddings(metadata=synthetic.make_metadata(cfg))
    model = model.to(dev)
    gcomm = None
    if world > 1 and args.graft_comm:  # the gradient all-reduce through the C ABI (graft_allreduce_f32, library-owned NCCL communicator)
        from recommendersystem_b200.comm import GraftComm
        gcomm = GraftComm(device=local)
    opt = FlatAdamW(model, lr=cfg["learning_rate"], graft_comm=gcomm)
    tw = make_task_weights(finetune_metric="rating") if args.config == "C3" else make_task_weights()
    S = cfg["max_sequence_length"]
    n_inter = rows * S

    pool = synthetic.make_batches(cfg, rows, 4, seed=1000 + rank, **synthetic.sampler_kwargs(args.config))
    host = [{k: torch.as_tensor(v).pin_memory() for k, v in b.items()} for b in pool]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

    from recommendersystem_b200.graphs import GraphedStep

    def eager_fwd_bwd(i, e2e):
        if e2e:
            d = {k: v.to(dev, non_blocking=True) for k, v in host[i % len(host)].items()}
        else:
            d = {k: v.clone() for k, v in resident[i % len(resident)].items()}  # the step edits its batch in place
        losses = model(d, False)
        loss = sum(l * w for l, w in zip(losses, tw))
        loss.backward()
        return torch.stack([l.detach() for l in losses])

    graphed = None
    if cfg.get("vocab_parallel") and world > 1:
        args.no_graph = True   # the vocab-parallel head exchanges rows through NCCL inside the step: launched eagerly, not captured
    if not args.no_graph:
        try:
            graphed = GraphedStep(model, tw, resident[0])
            opt.zero_grad()
        except Exception as e:  # capture is an optimisation: report and run eagerly
            sys.stderr.write(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graphed = None
            torch.cuda.synchronize()
            opt.zero_grad()

    def one_step(i, e2e, eager=False):
        if graphed is not None and not eager:
            losses = graphed(host[i % len(host)] if e2e else resident[i % len(resident)])
        else:
            losses = eager_fwd_bwd(i, e2e)
        opt.step()
        opt.zero_grad()
        if e2e:
            loss_host.copy_(losses, non_blocking=False)  # D2H read of the step's result
        return losses

    def timed(n_steps, e2e, profile=False):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        L.LAUNCHES = 0
        L.GEMM_EVENTS = [] if profile else None
        L.ATTN_EVENTS = [] if profile else None
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n_steps):
            one_step(i, e2e, eager=profile)  # per-GEMM events need eager launches
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        ev, L.GEMM_EVENTS = L.GEMM_EVENTS, None
        aev, L.ATTN_EVENTS = L.ATTN_EVENTS, None
        return float(ms) / n_steps, L.LAUNCHES, (ev, aev)

    for i in range(args.warmup):
        one_step(i, False)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, launches, _ = timed(args.steps, False)
    ms_e2e, _, _ = timed(args.steps, True)
    clocks = sampler.stop() if sampler else None
    # dominant-kernel roofline: the same steps launched eagerly with a CUDA event pair around every GEMM
    prof_steps = min(args.steps, 2)
    one_step(0, False, eager=True)  # untimed: the eager path re-populates the allocator pools the graph does not share
    ms_prof, _, (gemm_raw, attn_ev) = timed(prof_steps, False, profile=True)
    flags = L.status_flags(clear=True)  # preconditions the kernels check on the device (attention range, selection overflow, ...)

    out = None
    if rank == 0:
        pk = peaks()
        # FLOPs are booked from the rows a launch actually had to compute: the cloze-head GEMMs carry device-side row / depth
        # limits (the number of masked rows of the batch), read back here once; capacity FLOPs would overstate them ~4x
        gemm_ev = []
        for e0, e1, cap_flops, tag, mc, kc in gemm_raw:
            M, N, K = tag[0], tag[1], tag[2]
            if mc is not None:
                M = min(M, int(mc))
            if kc is not None:
                K = min(K, int(kc))
            gemm_ev.append((e0, e1, 2.0 * M * N * K, tag))
        flops = sum(e[2] for e in gemm_ev)
        gms = sum(e[0].elapsed_time(e[1]) for e in gemm_ev)
        # attention: 4 * d * sum over user runs of (tokens in the run)^2 per layer forward, 2.5x backward;
        # sum_t (seg_end[t] - seg_start[t]) over tokens = sum over runs of len^2
        attn = {"fwd": [0.0, 0.0, 0], "bwd": [0.0, 0.0, 0]}
        sq_cache = {}
        for e0, e1, kind, H, ss, se in attn_ev:
            key = ss.data_ptr()
            if key not in sq_cache:
                sq_cache[key] = float((se.long() - ss.long()).sum())
            fl = 4.0 * H * 64 * sq_cache[key] * (1.0 if kind == "fwd" else 2.5)
            a = attn[kind]; a[0] += fl; a[1] += e0.elapsed_time(e1); a[2] += 1
        attn_flops = attn["fwd"][0] + attn["bwd"][0]
        attn_ms = attn["fwd"][1] + attn["bwd"][1]
        if args.gemm_table:  # per-shape breakdown of the GEMM family (M is the launch's row capacity: masked-row GEMMs do less)
            agg = {}
            for e0, e1, f, tag in gemm_ev:
                a = agg.setdefault(tag, [0, 0.0, 0.0])
                a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += f
            EPI = ["bf16", "f32", "f32+=", "resid", "bias", "bias_gelu", "ce_partial", "ce_grad", "table", "swiglu", "rope", "bf16+=", "swiglu_bwd"]
            with open(args.gemm_table, "w") as fh:
                fh.write(f"GEMM launches of {prof_steps} eager C2 steps by shape (CUDA events around each launch; peak {pk['tensor']:.0f} TFLOP/s)\n\n")
                fh.write("| M | N | K | A | B | epilogue | split-K | launches | ms/step | TFLOP/s | frac | ms lost vs peak /step |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
                for tag, (c, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    tf = f / (ms * 1e-3) / 1e12
                    fh.write(f"| {tag[0]} | {tag[1]} | {tag[2]} | {'K' if tag[3] else 'MN'} | {'K' if tag[4] else 'MN'} | {EPI[tag[5]]} | {tag[6]} | {c // prof_steps} | "
                             f"{ms / prof_steps:.3f} | {tf:.0f} | {tf / pk['tensor']:.2f} | {(ms - f / (pk['tensor'] * 1e12) * 1e3) / prof_steps:.3f} |\n")
        ach = flops / (gms * 1e-3) / 1e12 if gms > 0 else 0.0

        def tf(fl, ms):
            return fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        out = {
            "metric": METRIC, "value": world * n_inter / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.config, world, graphed is not None),
            "e2e": {"value": world * n_inter / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (all trunk/head/grad GEMM launches of the timed steps)",
                         "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"], "traffic": (gemm_traffic() or {}).get("bytes"),
                         "traffic_note": (lambda t: None if t is None else f"DRAM read+write bytes of one launch ({t['launch']}), ncu --set full, {t['source']}")(gemm_traffic()),
                         "peak_source": pk["src"] + " sustained bf16", "launches": len(gemm_ev),
                         "share_of_step": gms / (ms_prof * prof_steps),
                         "measured": f"{prof_steps} extra eager steps ({ms_prof:.1f} ms/step) with an event pair around each GEMM launch; "
                                     "FLOPs booked from the device-side row counts of the head GEMMs (rows with weight > 0), not launch capacity"},
            # the second-largest kernel family, same method (event pair around every graft_attn_fwd / graft_attn_bwd call)
            "roofline_attention": {
                "bound": "tensor", "unit": "TFLOP/s", "peak": pk["tensor"],
                "flops": "4*d*sum(run tokens^2) per layer forward, 2.5x backward; masked / padded tiles not counted",
                "fwd": {"achieved": tf(attn["fwd"][0], attn["fwd"][1]), "frac": tf(attn["fwd"][0], attn["fwd"][1]) / pk["tensor"],
                        "launches": attn["fwd"][2], "us_per_launch": 1e3 * attn["fwd"][1] / max(1, attn["fwd"][2])},
                "bwd": {"achieved": tf(attn["bwd"][0], attn["bwd"][1]), "frac": tf(attn["bwd"][0], attn["bwd"][1]) / pk["tensor"],
                        "launches": attn["bwd"][2], "us_per_launch": 1e3 * attn["bwd"][1] / max(1, attn["bwd"][2])},
                "share_of_step": attn_ms / (ms_prof * prof_steps)},
            # whole step against the tensor peak: algorithmic FLOPs of every GEMM and attention launch of one step / step time
            "step_roofline": {"algorithmic_tflop_per_step": (flops + attn_flops) / prof_steps / 1e12,
                              "achieved": tf((flops + attn_flops) / prof_steps, ms_dev), "peak": pk["tensor"],
                              "frac": tf((flops + attn_flops) / prof_steps, ms_dev) / pk["tensor"], "unit": "TFLOP/s"},
            "status_flags": flags,
            "cuda_graph": graphed is not None,
            "clocks": clocks,
        }
    if world > 1:
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--cpu-rows", type=int, default=8, help="rows of the workload per CPU step (halved until the run fits --cpu-budget)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds the --impl reference run may take")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gemm-table", default=None, help="write a per-shape table of the GEMM launches (markdown) to this path")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--graft-comm", action="store_true", help="N > 1: all-reduce the gradient with graft_allreduce_f32 instead of torch.distributed")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU path on the host cores: EXACTLY --steps timed steps after --warmup warm-ups, each a bounded
        # sample of the workload (rows halved from 8 until the run fits ~4 minutes); same config / metric / unit as the GPU arm
        cb = cpu_reference(args.config, args.cpu_rows, args.steps, args.warmup, budget_s=args.cpu_budget)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"],
            "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.config, int(os.environ.get("WORLD_SIZE", "1")), not args.no_graph),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    out = gpu_bench(args)
    if rank == 0:
        if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args.config, args.cpu_rows, 2, 1, budget_s=40.0)
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out))


if __name__ == "__main__":
    main()

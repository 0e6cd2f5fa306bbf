"""Per-kernel roofline at C2 shapes, through the C ABI (recommendersystem_b200.lib).

For each hot kernel: algorithmic bytes or FLOPs per launch (no padding,
recompute or re-read credit) / median launch time from CUDA events on the launching
stream, against MEASURED_PEAKS.json.  L2 is flushed (256 MiB write) between launches.
Writes gpurun_out/kernel_roofline.{json,md}.

    python tools/kernel_roofline.py [--config C2] [--iters 7]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recommendersystem_b200 import lib as L, synthetic  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j.get("bf16_tflops_sustained", j["bf16_tflops"]), j["bf16_tflops"], "measured"
    return 6650.0, 1400.0, 1650.0, "fallback"


class Timer:
    def __init__(self, dev, iters):
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.iters = iters

    def __call__(self, fn, setup=None):
        ts = []
        for i in range(self.iters + 2):
            if setup is not None:
                setup()
            self.flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--iters", type=int, default=7)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L.load()
    hbm, tens, tens_burst, src = peaks()
    cfg, rows = synthetic.make_config(args.config)
    S, d, H, G, I = cfg["max_sequence_length"], cfg["embed_dim"], cfg["num_heads"], cfg["num_kv_heads"], cfg["intermediate_dim"]
    V0, V1 = cfg["vocab_sizes"]["0_matchedid"], cfg["vocab_sizes"]["1_matchedid"]
    V = V0 + V1
    N, T = rows * S, 2 * rows * S
    QKV = (H + 2 * G) * 64
    timer = Timer(dev, args.iters)
    out = []

    def rec(name, ms, *, bytes_=None, flops=None, note=""):
        if bytes_ is not None:
            ach = bytes_ / (ms * 1e-3) / 1e9
            out.append(dict(kernel=name, bound="hbm", ms=ms, algorithmic=bytes_, achieved=ach, unit="GB/s", peak=hbm,
                            frac=ach / hbm, frac_nominal_8TBs=ach / 8000.0, note=note))
        else:
            ach = flops / (ms * 1e-3) / 1e12
            out.append(dict(kernel=name, bound="tensor", ms=ms, algorithmic=flops, achieved=ach, unit="TFLOP/s", peak=tens,
                            frac=ach / tens, frac_burst=ach / tens_burst, note=note))
        print(f"{name:44s} {ms * 1e3:9.1f} us  {out[-1]['achieved']:9.1f} {out[-1]['unit']:8s} {100 * out[-1]['frac']:5.1f}%  {note}",
              flush=True)

    g = torch.Generator(device=dev).manual_seed(0)
    batch = synthetic.make_batches(cfg, rows, 1, seed=1000)[0]

    # ---------------------------------------------------------------- gather (HBM) ----
    table = torch.randn((V + 1, d), device=dev, generator=g)
    idx = torch.as_tensor(batch["matchedid"]).to(dev)
    x0 = torch.empty((T, d), device=dev)
    ms = timer(lambda: L.item_gather(table, idx, x0))
    rec("item_gather (fp32 rows)", ms, bytes_=N * (4 + 2 * d * 4), note=f"N={N} d={d} V={V}")

    # ------------------------------------------------------------ scatter-add (HBM) ----
    def scatter_case(n, dd, zipf, tag):
        smp = synthetic.HistorySampler(cfg, seed=7, zipf_s=zipf)
        ids = torch.as_tensor(smp._perm[np.searchsorted(smp._cdf, smp.rng.random(n))].astype(np.int32)).to(dev)
        gsrc = torch.randn((n, dd), device=dev, generator=g)
        gt = torch.zeros((V + 1, dd), device=dev)
        uniq = int(torch.unique(ids).numel())
        ms = timer(lambda: L.scatter_add_sorted(gsrc, dd, ids, dd, gt))
        rec(f"scatter_add_sorted {tag}", ms, bytes_=n * (4 + dd * 4) + uniq * dd * 8,
            note=f"n={n} d={dd} unique_rows={uniq} zipf={zipf} (incl. radix sort + fix-up)")

    scatter_case(N, d, 1.0, "C2 step")
    scatter_case(N, d, 1.2, "C5 n=64k d=768")
    scatter_case(1 << 20, d, 1.2, "C5 n=1M d=768")
    scatter_case(N, 2048, 1.2, "C5 n=64k d=2048")
    scatter_case(1 << 20, 2048, 1.2, "C5 n=1M d=2048")

    # ------------------------------------------------------------------ norms (HBM) ----
    x = torch.randn((T, d), device=dev, generator=g)
    scale = torch.ones(d, device=dev)
    y16 = torch.empty((T, d), dtype=torch.bfloat16, device=dev)
    rstd = torch.empty(T, device=dev)
    ms = timer(lambda: L.rmsnorm_fwd(x, scale, y16, None, rstd))
    rec("rmsnorm_fwd", ms, bytes_=T * d * (4 + 2) + T * 4)
    dlt = torch.randn((T, d), device=dev, generator=g).bfloat16()
    hsum = torch.empty((T, d), device=dev)
    ms = timer(lambda: L.rmsnorm_fwd(x, scale, y16, None, rstd, delta=dlt, h_out=hsum))
    rec("rmsnorm_fwd + fused residual add", ms, bytes_=T * d * (4 + 2 + 4 + 2) + T * 4)
    del dlt, hsum
    dy16 = torch.randn((T, d), device=dev, generator=g).bfloat16()
    dres = torch.randn((T, d), device=dev, generator=g)
    dx = torch.empty((T, d), device=dev)
    dx16 = torch.empty((T, d), dtype=torch.bfloat16, device=dev)
    dsc = torch.empty(d, device=dev)
    ms = timer(lambda: L.rmsnorm_bwd(x, scale, rstd, dy16, None, dres, dx, dx16, dsc))
    rec("rmsnorm_bwd (+residual grad, +bf16 copy)", ms, bytes_=T * d * (4 + 2 + 4 + 4 + 2))
    gu = torch.randn((T, 2 * I), device=dev, generator=g).bfloat16()
    act = torch.empty((T, I), dtype=torch.bfloat16, device=dev)
    ms = timer(lambda: L.swiglu_fwd(gu, act))
    rec("swiglu_fwd", ms, bytes_=T * I * 2 * 3)
    dgu = torch.empty_like(gu)
    ms = timer(lambda: L.swiglu_bwd(gu, act, dgu))
    rec("swiglu_bwd", ms, bytes_=T * I * 2 * 5)

    # ------------------------------------------------------------- attention (tensor) ----
    uid = torch.as_tensor(batch["userid"]).to(dev).reshape(rows, S)
    u = batch["userid"].reshape(rows, S)
    pairs = 0
    for r in u:
        ch = np.flatnonzero(r[1:] != r[:-1]) + 1
        ls = np.diff(np.concatenate([[0], ch, [len(r)]]))
        pairs += int(((2 * ls) ** 2).sum())
    attn_flops = 4.0 * H * 64 * pairs
    qkv = (torch.randn((T, QKV), device=dev, generator=g) * 0.5).bfloat16()
    # mask ids as a training step has them after mask_tokens (id 1 on ~10 % of the positions; round 1 timed all-zero ids,
    # which skips the per-element select path of every 8-column group that holds a masked key)
    tm = (torch.rand(T // 2, device=dev, generator=g) < 0.1).to(torch.int32).repeat_interleave(2).contiguous()
    seg_s, seg_e = L.token_segments(uid)
    at = L.attn_args(qkv, H, G, rows, 2 * S, tm, seg_s, seg_e)
    ao = torch.empty((T, H * 64), dtype=torch.bfloat16, device=dev)
    lse = torch.empty((H, T), device=dev)
    ms = timer(lambda: L.attn_fwd(at, ao, lse))
    rec("attention fwd (jagged, GQA pair)", ms, flops=attn_flops, note=f"sum l_tok^2={pairs} ({pairs / (rows * (2 * S) ** 2):.2f} of dense); mask id 1 on 10 % of the positions")
    dao = torch.randn((T, H * 64), device=dev, generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((H, T), device=dev)
    ms = timer(lambda: L.attn_bwd(at, ao, dao, lse, dqkv, delta))
    rec("attention bwd (delta + dq + dkv)", ms, flops=2.5 * attn_flops)

    # one user per row: every query tile visits all 16 key tiles (twice the tile steps of the jagged batch with the
    # same number of CTAs) - separates the per-CTA fixed cost from the per-tile cost
    uid1 = torch.arange(1, rows + 1, dtype=torch.int32, device=dev)[:, None].expand(rows, S).contiguous()
    s1, e1 = L.token_segments(uid1)
    at1 = L.attn_args(qkv, H, G, rows, 2 * S, tm, s1, e1)
    dense_flops = 4.0 * H * 64 * rows * (2 * S) ** 2
    ms = timer(lambda: L.attn_fwd(at1, ao, lse))
    rec("attention fwd, one user per row (dense)", ms, flops=dense_flops, note="16 key tiles per query tile")
    ms = timer(lambda: L.attn_bwd(at1, ao, dao, lse, dqkv, delta))
    rec("attention bwd, one user per row (dense)", ms, flops=2.5 * dense_flops)

    # ------------------------------------------------------------------ GEMMs (tensor) ----
    def gemm_case(name, M, Nn, K, **kw):
        A = (torch.randn((M, K), device=dev, generator=g) * 0.1).bfloat16()
        B = (torch.randn((Nn, K), device=dev, generator=g) * 0.1).bfloat16()
        C = torch.empty((M, Nn), dtype=torch.bfloat16, device=dev)
        ms = timer(lambda: L.gemm(A, B, C, M, Nn, K, **kw))
        rec(name, ms, flops=2.0 * M * Nn * K, note=f"{M}x{Nn}x{K}")

    gemm_case("gemm qkv proj", T, QKV, d)
    gemm_case("gemm w1|w3 (gate,up)", T, 2 * I, d)
    # the fused epilogues the step actually uses
    A = (torch.randn((T, d), device=dev, generator=g) * 0.1).bfloat16()
    Wg = (torch.randn((2 * I, d), device=dev, generator=g) * 0.1).bfloat16()
    gu_f = torch.empty((T, 2 * I), dtype=torch.bfloat16, device=dev); act_f = torch.empty((T, I), dtype=torch.bfloat16, device=dev)
    ms = timer(lambda: L.gemm(A, Wg, gu_f, T, 2 * I, d, epilogue=L.EPI_SWIGLU, aux=act_f, ldaux=I))
    rec("gemm w1|w3 + fused SwiGLU epilogue", ms, flops=2.0 * T * 2 * I * d, note=f"{T}x{2 * I}x{d}, writes gu and act")
    Wq = (torch.randn((QKV, d), device=dev, generator=g) * 0.1).bfloat16()
    qkv_f = torch.empty((T, QKV), dtype=torch.bfloat16, device=dev)
    inv = 1.0 / (500000.0 ** (torch.arange(0, 64, 2).float() / 64))
    ang = torch.outer(torch.arange(2 * S).float(), inv)
    cosb, sinb = torch.cos(ang).to(dev).contiguous(), torch.sin(ang).to(dev).contiguous()
    ms = timer(lambda: L.gemm(A, Wq, qkv_f, T, QKV, d, epilogue=L.EPI_ROPE, rope=dict(cos=cosb, sin=sinb, cols=(H + G) * 64, row_len=2 * S)))
    rec("gemm qkv proj + fused RoPE epilogue", ms, flops=2.0 * T * QKV * d, note=f"{T}x{QKV}x{d}")
    cost, sint = cosb.t().contiguous(), sinb.t().contiguous()
    ms = timer(lambda: L.gemm(A, Wq, qkv_f, T, QKV, d, epilogue=L.EPI_ROPE, rope=dict(cos=cost, sin=sint, t_stride=2 * S, cols=(H + G) * 64, row_len=2 * S)))
    rec("gemm qkv proj + fused RoPE epilogue, transposed tables (step path)", ms, flops=2.0 * T * QKV * d, note=f"{T}x{QKV}x{d}")
    del A, Wg, gu_f, act_f, Wq, qkv_f
    # d(act) = dx . W2 with the SwiGLU backward fused into the epilogue (reads gu, writes dgu; d(act) stays on chip)
    dxa = (torch.randn((T, d), device=dev, generator=g) * 0.5).bfloat16()
    w2m = (torch.randn((d, I), device=dev, generator=g) * 0.05).bfloat16()
    guu = torch.randn((T, 2 * I), device=dev, generator=g).bfloat16()
    dguu = torch.empty_like(guu)
    ms = timer(lambda: L.gemm(dxa, w2m, dguu, T, I, d, b_kmajor=False, ldb=I, ldc=2 * I, epilogue=L.EPI_SWIGLU_BWD, aux=guu, ldaux=2 * I))
    rec("gemm d(act) + fused SwiGLU backward epilogue", ms, flops=2.0 * T * I * d,
        note=f"{T}x{I}x{d}; also HBM-heavy: {(2 * T * 2 * I * 2 + T * d * 2) / (ms * 1e-3) / 1e9:.0f} GB/s of gu + dgu + dx")
    del dxa, w2m, guu, dguu
    gemm_case("gemm w2 (down)", T, d, I)
    gemm_case("gemm out proj", T, d, H * 64)
    # dW shape: [2I, T] x [T, d], MN-major operands, split-K
    A = (torch.randn((T, 2 * I), device=dev, generator=g) * 0.1).bfloat16()
    B = (torch.randn((T, d), device=dev, generator=g) * 0.1).bfloat16()
    C = torch.empty((2 * I, d), device=dev)
    ms = timer(lambda: L.gemm(A, B, C, 2 * I, d, T, a_kmajor=False, lda=2 * I, b_kmajor=False, ldb=d, epilogue=L.EPI_F32, split_k=-1))
    rec("gemm dW w1|w3 (MN-major, split-K)", ms, flops=2.0 * 2 * I * d * T, note=f"{2 * I}x{d}x{T}")

    # ------------------------------------------------- catalog logits + CE (tensor) ----
    cap = cfg["mask_topk"] * rows
    hs = (torch.randn((cap, d), device=dev, generator=g) * 0.5).bfloat16()
    tab16 = (torch.randn((V0, d), device=dev, generator=g) * 0.05).bfloat16()
    nb = L.ce_nblocks(V0)
    pmax = torch.empty((nb, cap), device=dev); psum = torch.empty((nb, cap), device=dev)
    tgt = torch.zeros(cap, device=dev)
    pos = torch.randint(0, V0, (cap,), device=dev, dtype=torch.int32, generator=g)
    ms = timer(lambda: L.gemm(hs, tab16, None, cap, V0, d, epilogue=L.EPI_CE_PARTIAL, ce=dict(pmax=pmax, psum=psum, tgt=tgt, target=pos)))
    rec("catalog logits + fused CE fwd", ms, flops=2.0 * cap * V0 * d, note=f"{cap} rows x {V0} items; logits never stored "
        f"(would be {cap * V0 * 2 / 1e9:.2f} GB bf16)")
    lse_r = torch.full((cap,), 11.0, device=dev); coef = torch.full((cap,), 1e-3, device=dev)
    CHK = 16384
    dlog = torch.empty((cap, CHK), dtype=torch.bfloat16, device=dev)
    ms = timer(lambda: L.gemm(hs, tab16[:CHK], dlog, cap, CHK, d, ldc=CHK, epilogue=L.EPI_CE_GRAD, ce=dict(lse=lse_r, coef=coef, target=pos)))
    rec("catalog logit gradient chunk (CE_GRAD epilogue)", ms, flops=2.0 * cap * CHK * d, note=f"{cap} rows x {CHK} items, all rows live")

    # ---------------------------------------------------------------- optimiser (HBM) ----
    n = 135_000_000
    p_, g_, m_, v_ = (torch.zeros(n, device=dev) for _ in range(4))
    sc = torch.ones(1, device=dev)
    ms = timer(lambda: L.adamw(p_, g_, m_, v_, 1e-4, 0.9, 0.95, 1e-8, 0.1, 1, sc))
    rec("adamw (flat fp32)", ms, bytes_=n * 4 * 7, note=f"{n} params")

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    meta = dict(config=args.config, peaks=dict(hbm_gbs=hbm, bf16_tflops_sustained=tens, bf16_tflops_burst=tens_burst, source=src),
                iters=args.iters, l2="256 MiB flush write before every timed launch", kernels=out)
    json.dump(meta, open(os.path.join(ROOT, "gpurun_out", "kernel_roofline.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_roofline.md"), "w") as f:
        f.write(f"# Per-kernel roofline, {args.config} shapes (CUDA events, median of {args.iters}, L2 flushed)\n\n")
        f.write(f"Peaks ({src}): HBM {hbm} GB/s, bf16 {tens} TFLOP/s sustained ({tens_burst} burst).\n\n")
        f.write("| kernel | us | achieved | unit | % of peak | note |\n|---|---|---|---|---|---|\n")
        for k in out:
            f.write(f"| {k['kernel']} | {k['ms'] * 1e3:.1f} | {k['achieved']:.1f} | {k['unit']} | {100 * k['frac']:.1f} | {k['note']} |\n")


if __name__ == "__main__":
    main()
